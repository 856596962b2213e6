#!/usr/bin/env python
"""bench.py — 608x608 images/sec of the YOLOv4 hot path (110-conv forward + head decode + per-class NMS).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 608] [--batch 32]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling)

One "step" = one pass of the hot path over one batch (BASELINE.json configs[1]: batch 32, 608x608, fp16 tensor
cores, 1xB200).  `value` times K steps with the inputs resident in HBM (CUDA events on the engine stream, max
over ranks); `e2e` times the same metric through the C-ABI call y4_predict() with pinned HOST buffers (H2D of
the images and D2H of the detections inside the timed region).  `roofline` is the conv stack (tcgen05 kernels):
algorithmic conv FLOPs per step / its CUDA-event duration, against the measured dense-bf16 peak.
`--impl reference` times a CPU port of the reference (oracle/y4_cpu_fast.py: torch-CPU / oneDNN convs, numpy decode,
C + OpenMP NMS, all host threads) on a bounded sample of the same workload: TensorFlow is not installable in this image,
so the reference's own tf.keras CPU forward cannot be run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = '608x608 images/sec'


def conv_gflop(size):
    import netspec
    return netspec.conv_gflop(size)


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) when present, else the B200_PROFILING.md fallback.
    Both dense-bf16 figures are kept: `burst` (a kernel timed alone, SM clocks near max) and `sustained` (a seconds-long
    loop under the 1 kW power cap); the caller picks by the SM clock it saw and reports the fraction against both."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        burst = float(p.get('bf16_tflops', p.get('bf16_tflops_sustained')))
        return {'tflops_burst': burst, 'tflops_sustained': float(p.get('bf16_tflops_sustained', burst)), 'hbm_gbs': float(p['hbm_gbs']),
                'source': 'MEASURED_PEAKS.json'}
    return {'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'hbm_gbs': 6650.0,
            'source': 'fallback (B200_PROFILING.md: 1.59 PF burst / 1.4 PF sustained, 6.65 TB/s)'}


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a thread every
    5 ms (a 20-step timed region lasts ~150 ms: `nvidia-smi -lms 100` caught one sample); nvidia-smi is the fallback."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.rows, self.proc, self.nvml, self.stop_flag = device, [], None, None, False
        self.marks = {}

    def mark(self, name):
        self.marks[name] = time.time()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[self.device]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else self.device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {'hw_slowdown': n.nvmlClocksEventReasonHwSlowdown if hasattr(n, 'nvmlClocksEventReasonHwSlowdown') else 0x8,
                'hw_thermal_slowdown': getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                'sw_thermal_slowdown': getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                'sw_power_cap': getattr(n, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or getattr(n, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                r = int(get_reasons(self.h))
                self.rows.append((time.time(), mhz, sorted(k for k, b in bits.items() if r & b)))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(',')]
            if len(c) >= 7 and c[0].replace('.', '').isdigit():
                self.max_mhz = float(c[1]) if c[1].replace('.', '').isdigit() else None
                self.rows.append((time.time(), float(c[0]), [names[i] for i in range(4) if c[3 + i].lower().startswith('active')]))

    def summary(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        sm = [r[1] for r in rows]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': getattr(self, 'max_mhz', None),
                'reasons': sorted({x for r in rows for x in r[2]}), 'samples': len(sm)}

    def stop(self):
        """Summary over the TIMED region (marks 'timed0'..'timed1'); the whole loaded span is reported beside it."""
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no clock samples (NVML and nvidia-smi unavailable)'], 'samples': 0}
        m = self.marks
        out = self.summary(m.get('timed0', 0.0), m.get('timed1', 1e30))
        load = self.summary(m.get('load0', 0.0), m.get('load1', 1e30))
        out['under_load'] = {'sm_mhz': load['sm_mhz'], 'samples': load['samples'], 'reasons': load['reasons'],
                             'span': 'first warm-up step .. last profiled launch'}
        out['source'] = 'NVML polled every 5 ms' if self.nvml else 'nvidia-smi -lms 20'
        if out['samples'] < 10:                       # a very short timed region: fall back to the loaded span, and say so
            out.update(sm_mhz=load['sm_mhz'], reasons=load['reasons'], samples=load['samples'], window='under_load (timed region held < 10 samples)')
        return out


def cpu_numpy_images_per_sec(size, n_images, weights):
    """The numpy oracle (OpenBLAS sgemm convs, Python-loop NMS) on a bounded sample."""
    import y4_oracle as O
    imgs = O.synth_images(0, 0, n_images, size)
    t0 = time.perf_counter()
    O.predict(imgs, weights)
    dt = time.perf_counter() - t0
    return n_images / dt, dt


class CpuReference:
    """The reference's path on the host cores, best CPU kernels available in this image: torch CPU convs (oneDNN, fp32,
    channels_last, all threads) + numpy decode + C/OpenMP combined NMS (oracle/y4_cpu_fast.py).  The reference's own
    tf.keras forward cannot run here: TensorFlow is not installed and there is no network; if a later image
    ships it, `tensorflow` shows up in `kind_note`."""

    def __init__(self, weights):
        import y4_cpu_fast as F
        self.net = F.TorchNet(weights)
        try:
            import tensorflow as tf       # noqa: F401  (never present in this image; recorded if it ever is)
            self.tf = tf.__version__
        except Exception:
            self.tf = None
        import torch
        self.threads = torch.get_num_threads()

    def images_per_sec(self, size, n_images, first=0):
        import y4_oracle as O
        imgs = O.synth_images(0, first, n_images, size)
        t0 = time.perf_counter()
        out = self.net.predict(imgs)
        dt = time.perf_counter() - t0
        return n_images / dt, dt, out

    def describe(self):
        return ('torch CPU (oneDNN) fp32 convs + numpy decode + C/OpenMP NMS, same op order as the reference; '
                + (f'tensorflow {self.tf} is importable but /root/reference is not on this box' if self.tf
                   else 'tf.keras itself is not installable in this image'))


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores (CpuReference; kind "port")."""
    if rank != 0:
        return
    import y4_oracle as O
    cores = len(os.sched_getaffinity(0))
    W = O.synth_weights(seed=1)
    ref = CpuReference(W)
    _, t1, _ = ref.images_per_sec(args.size, 1)                        # untimed calibration / warm-up (oneDNN primitive caches)
    # a step = a bounded sample of the batch, sized so that K steps + W warm-ups end within a few minutes
    sample = int(max(1, min(args.batch, 150.0 / max(t1, 1e-3) / max(args.steps + args.warmup, 1))))
    for _ in range(args.warmup):
        ref.images_per_sec(args.size, sample)
    t0 = time.perf_counter()
    for k in range(args.steps):
        ref.images_per_sec(args.size, sample, first=k * sample)
    dt = time.perf_counter() - t0
    v = args.steps * sample / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'configs[1]: yolov4 forward + decode + NMS {args.size}x{args.size}, 80 classes; {sample} image(s) per step '
                                   f'(bounded sample of the batch of {args.batch})'},
            'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'threads': ref.threads, 'kind': 'port',
                             'sample': f'{sample} image(s) per step x {args.steps} steps; ' + ref.describe()},
            'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """Native libraries (NCCL prints its version banner) write to fd 1; the driver expects ONE JSON line there.  Point fd 1 at
    stderr for the whole run and keep a private duplicate for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=608)
    ap.add_argument('--batch', type=int, default=32, help='images per GPU per step')
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'fp32', 'fp16x3'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity-modes', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    dist = None
    if world > 1:
        import torch.distributed as dist       # host-side rendezvous only (uid broadcast, barrier, max-reduce)
        dist.init_process_group('gloo', rank=rank, world_size=world)

    import y4b200
    import y4_oracle as O
    from y4b200 import binding

    S, B = args.size, args.batch
    prec = {'fp16': y4b200.PREC_FP16, 'fp32': y4b200.PREC_FP32, 'fp16x3': y4b200.PREC_FP16X3}[args.precision]
    eng = y4b200.Engine(img_size=S, max_batch=B, precision=prec, device=local)
    W = O.synth_weights(seed=1)
    eng.load_darknet_bytes(W.to_darknet_bytes())
    if world > 1:
        import torch
        uid = torch.from_numpy(eng.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(rank, world, uid.numpy())

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()

    def step_resident():
        eng.run_resident(B)
        if world > 1:
            eng.allgather_results(B, fetch=False)

    # ---- resident timing: value -------------------------------------------------------------------
    clocks = ClockSampler(local)                       # nvidia-smi takes a few hundred ms to produce its first line: start it early,
    clocks.start()                                     # keep only the samples taken between the first warm-up step and the last timed launch
    eng.synth_fill(0, rank * B, B)                     # image i depends only on its global index
    clocks.mark('load0')
    for _ in range(args.warmup):
        step_resident()
    barrier()
    l0 = eng.launch_count()
    clocks.mark('timed0')
    eng.timer_begin()
    for _ in range(args.steps):
        step_resident()
    ms = eng.timer_end()
    clocks.mark('timed1')
    launches = eng.launch_count() - l0
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- conv stack alone, and the dominant kernel: roofline numerators ----------------------------
    eng.timer_begin()
    for _ in range(args.steps):
        eng.run_forward_resident(B)
    fwd_ms = eng.timer_end() / args.steps
    eng.timer_begin()
    for _ in range(args.steps):
        eng.run_decode_nms_resident(B)
    dn_ms = eng.timer_end() / args.steps
    # per-launch durations (CUDA events on the engine stream around every launch of the forward), grouped by kernel
    # instantiation: the instantiation with the largest share of the step is the one the roofline object describes
    prof = np.median(np.stack([eng.profile_layers(B) for _ in range(5)]), axis=0)
    eng.sync()
    clocks.mark('load1')
    clk = clocks.stop()
    pk = peaks()
    # denominator: the burst figure when the SM clock stayed near its maximum during the timed region (a ~150 ms region does
    # not reach the power cap), the sustained one otherwise; the fraction against BOTH is reported
    near_max = bool(clk.get('sm_mhz') and clk.get('sm_max_mhz') and clk['sm_mhz'] >= 0.9 * clk['sm_max_mhz'])
    peak_kind = 'burst' if near_max else 'sustained'
    peak_tf = pk['tflops_burst'] if near_max else pk['tflops_sustained']
    groups = {}
    for t, l in zip(prof, eng.steps()):                # one entry per launch of the forward, in schedule order
        if l['kernel_kind'] in (1, 2):
            nepi, lean = (4, 1) if l['tc_epi_warps'] == 44 else (l['tc_epi_warps'], 0)
            name = (f"conv_tc2_kernel<{l['tile_n']}, {nepi}>" if l['tc_mode'] in (4, 5)
                    else f"conv_tc_kernel<{l['tile_n']}, {l['tc_bk']}, {1 if args.precision == 'fp16x3' else 0}, {nepi}, {lean}>")
        else:
            name = {5: 'spp_sep_kernel', 4: 'conv0_tc_kernel', 3: 'conv0_direct_kernel', 0: 'conv_simt_kernel'}.get(l['kernel_kind'], 'other')
        g = groups.setdefault(name, [0, 0.0, 0.0])
        g[0] += 1; g[1] += float(t); g[2] += l['flops'] * B
    top = max(groups, key=lambda k: groups[k][1])
    n_top, ms_top, fl_top = groups[top]
    achieved = fl_top / ms_top / 1e9                   # FLOP / ms / 1e9 == TFLOP/s
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if top in tj.get('kernels', {}):
            traffic = tj['kernels'][top]['dram_bytes_per_launch']
    gflop_step = conv_gflop(S) * B
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                'frac_of_burst_peak': achieved / pk['tflops_burst'], 'frac_of_sustained_peak': achieved / pk['tflops_sustained'],
                'peak_kind': f"{peak_kind} (median SM clock {clk.get('sm_mhz')} MHz of {clk.get('sm_max_mhz')} during the timed region)",
                'traffic': traffic, 'kernel': top, 'launches_per_step': n_top, 'avg_launch_us': 1e3 * ms_top / n_top,
                'algorithmic_flops_per_launch': fl_top / n_top, 'share_of_forward': ms_top / float(prof.sum()),
                'traffic_source': 'profiles/traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of this kernel, from the committed capture (not re-measured in this run)' if traffic else None,
                'peak_source': pk['source'],
                'conv_stack': {'achieved': gflop_step / fwd_ms, 'frac': gflop_step / fwd_ms / peak_tf, 'unit': 'TFLOP/s',
                               'frac_of_burst_peak': gflop_step / fwd_ms / pk['tflops_burst'],
                               'frac_of_sustained_peak': gflop_step / fwd_ms / pk['tflops_sustained'],
                               'flops_per_step': gflop_step * 1e9, 'forward_ms_per_step': fwd_ms,
                               'what': 'all 110 convs + SPP of one step (graph replay), algorithmic conv FLOPs / CUDA-event time'}}
    # decode + NMS (BASELINE metric 'decode+NMS us/img'): HBM bound; algorithmic bytes = the fp32 head tensors read once
    # (N * 85 * 4 B per image) + the 2,404 B result record; heads of one batch (247 MB) exceed the 126 MB L2
    dn_bytes = B * (eng.num_boxes * (5 + eng.num_classes) * 4 + eng.max_boxes * 24 + 4)
    decode_nms = {'bound': 'hbm', 'us_per_img': 1e3 * dn_ms / B, 'ms_per_batch': dn_ms, 'algorithmic_bytes_per_batch': dn_bytes,
                  'achieved': dn_bytes / dn_ms / 1e6, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': dn_bytes / dn_ms / 1e6 / pk['hbm_gbs'],
                  'kernels': 'decode_filter + nms_image + nms_overflow + nms_merge (graph replay, CUDA events)'}

    # ---- e2e through the C-ABI with host buffers --------------------------------------------------
    imgs = binding.pinned_array((B, S, S, 3))
    imgs[...] = O.synth_images(0, rank * B, B, S)
    e2e_steps = max(3, args.steps)
    imgs2 = binding.pinned_array((B, S, S, 3))
    imgs2[...] = imgs
    eng.predict(imgs)
    barrier()
    # blocking call (what Yolov4.predict_img does): H2D, compute and D2H strictly one after the other
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = eng.predict(imgs)
    barrier()
    e2e_sync_dt = time.perf_counter() - t0
    # pipelined call (what a batched caller like export_prediction does): H2D of batch i+1 overlaps compute of batch i
    bufs = [imgs, imgs2]
    for i in range(3):                                 # untimed: copy stream, second input slot, both graphs get created here
        eng.submit(bufs[i & 1])
        eng.collect()
    barrier()
    t0 = time.perf_counter()
    eng.submit(bufs[0])
    for i in range(1, e2e_steps):
        eng.submit(bufs[i & 1])
        out = eng.collect()
    out = eng.collect()
    barrier()
    e2e_dt = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    # raw 8-bit images, the form the reference's callers hold (cv2.imread -> preprocess_img -> predict, models.py:95-127,141-179):
    # H2D of the uint8 pixels, resize + /255 on the GPU, forward, decode, NMS, D2H of the detections
    raws = [binding.pinned_array((S, S, 3), np.uint8) for _ in range(2 * B)]
    rng = np.random.default_rng(1234 + rank)
    for r in raws:
        r[...] = rng.integers(0, 256, (S, S, 3), dtype=np.uint8)
    rb = [raws[:B], raws[B:]]
    for i in range(3):
        eng.submit_u8(rb[i & 1])
        eng.collect()
    barrier()
    t0 = time.perf_counter()
    eng.submit_u8(rb[0])
    for i in range(1, e2e_steps):
        eng.submit_u8(rb[i & 1])
        out8 = eng.collect()
    out8 = eng.collect()
    barrier()
    e2e8_dt = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e8_dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e8_dt = float(t.item())
    mb = eng.max_boxes
    d2h = int(B * (mb * 4 * 4 + mb * 4 * 3 + 4))
    e2e = {'value': world * B * e2e_steps / e2e8_dt, 'unit': 'images/s', 'h2d_bytes_per_step': int(B * S * S * 3),
           'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
           'api': 'y4_submit_u8/y4_collect, depth 2: raw uint8 HWC images in pinned host memory -> GPU resize (cv2 INTER_LINEAR semantics) + /255 '
                  '-> forward -> decode -> NMS -> detections in host memory (what Yolov4.export_prediction calls)',
           'float32_input': {'value': world * B * e2e_steps / e2e_dt, 'h2d_bytes_per_step': int(imgs.nbytes),
                             'api': 'y4_submit/y4_collect with already preprocessed float32 NHWC images'},
           'blocking_call_value': world * B * e2e_steps / e2e_sync_dt, 'blocking_call_api': 'y4_predict (float32 input, no overlap)'}

    if rank != 0:
        return
    line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp16': 'f16', 'fp32': 'f32', 'fp16x3': 'f16x3 (hi+lo fp16 operands, fp32 accumulate)'}[args.precision], 'data': 'synthetic',
            'config': {'workload': f'configs[1]: yolov4 CSPDarknet53+SPP+PANet forward + 3-scale decode + per-class NMS, '
                                   f'batch {B}/GPU, {S}x{S}, 80 classes, seeded random weights/images',
                       'l2': 'activations written+read per step (~7.3 GB at batch 32) >> 126 MB L2; no flush needed',
                       'global_batch': world * B, 'parallelism': f'dp{world}'},
            'gpu_launches': launches, 'clocks': clk, 'e2e': e2e, 'roofline': roofline,
            'decode_nms': decode_nms, 'decode_nms_us_per_img': 1e3 * dn_ms / B, 'valid_detections_img0': int(out[3][0])}
    eng.close()
    if world == 1 and not args.no_parity_modes:
        # the modes that meet the reference's fp32 results to the north-star tolerance (tests/test_gpu_split.py,
        # tests/test_gpu_forward.py), timed the same way as `value` (resident inputs, CUDA events, whole step)
        line['parity_modes'] = {'unit': 'images/s', 'how': 'same step and timing as `value`; fewer steps',
                                'tolerance': 'bit-exact indices/classes on round-off-stable images, boxes/scores <= max(1e-4, 3x oracle noise)'}
        for name, code, k in (('fp16x3', y4b200.PREC_FP16X3, max(3, args.steps // 4)), ('fp32', y4b200.PREC_FP32, 2)):
            if name == args.precision:
                line['parity_modes'][name] = value
                continue
            e2 = y4b200.Engine(img_size=S, max_batch=B, precision=code, device=local)
            e2.load_darknet_bytes(W.to_darknet_bytes())
            e2.synth_fill(0, rank * B, B)
            for _ in range(2 if name == 'fp32' else 3):
                e2.run_resident(B)
            e2.sync()
            e2.timer_begin()
            for _ in range(k):
                e2.run_resident(B)
            line['parity_modes'][name] = B * k / (e2.timer_end() * 1e-3)
            e2.close()
    if not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        ref = CpuReference(W)
        _, t1, _ = ref.images_per_sec(S, 1)                                # warm-up + calibration
        n = int(max(2, min(B, 15.0 / max(t1, 1e-3))))                      # ~15 s of CPU work
        v, dt, _ = ref.images_per_sec(S, n)
        vn, dtn = cpu_numpy_images_per_sec(S, 1, W)
        line['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'threads': ref.threads, 'kind': 'port',
                                'sample': f'{n} images {S}x{S} in one batch ({dt:.1f} s): ' + ref.describe(),
                                'numpy_port_value': vn, 'numpy_port_sample': f'1 image through the numpy + OpenBLAS oracle ({dtn:.1f} s)'}
    emit(line)


if __name__ == '__main__':
    main()
