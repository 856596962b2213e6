#!/usr/bin/env python
"""bench.py — 608x608 images/sec of the YOLOv4 hot path (110-conv forward + head decode + per-class NMS).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 608] [--batch 32]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling)

One "step" = one pass of the hot path over one batch (BASELINE.json configs[1]: batch 32, 608x608, fp16 tensor
cores, 1xB200).  `value` times K steps with the inputs resident in HBM (CUDA events on the engine stream, max
over ranks); `e2e` times the same metric through the C-ABI call y4_predict() with pinned HOST buffers (H2D of
the images and D2H of the detections inside the timed region).  `roofline` is the conv stack (tcgen05 kernels):
algorithmic conv FLOPs per step / its CUDA-event duration, against the measured dense-bf16 peak.
`--impl reference` times the CPU restatement of the reference (oracle/, numpy + BLAS, all host threads) on a
bounded sample of the same workload: TensorFlow is not installable in this image, so the reference's own
tf.keras CPU forward cannot be run.
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
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'tflops': float(p.get('bf16_tflops_sustained', p.get('bf16_tflops'))), 'hbm_gbs': float(p['hbm_gbs']),
                'source': 'MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)'}
    return {'tflops': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md: 1.4 PF sustained, 6.65 TB/s)'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(',')])

    def window(self, t0, t1):
        """Keep the samples taken while the GPU was under load ([t0, t1]); all of them if the window caught none."""
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        self.rows = inside or self.rows

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.window(getattr(self, 't0', 0.0), getattr(self, 't1', 1e30))
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r[1:] for r in self.rows]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def cpu_port_images_per_sec(size, n_images, weights):
    """The oracle (CPU restatement of the reference, numpy + multithreaded BLAS) on a bounded sample."""
    import y4_oracle as O
    imgs = O.synth_images(0, 0, n_images, size)
    t0 = time.perf_counter()
    O.predict(imgs, weights)
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; TF unavailable here)."""
    if rank != 0:
        return
    import y4_oracle as O
    cores = len(os.sched_getaffinity(0))
    W = O.synth_weights(seed=1)
    sample = 1                                   # images per step: bounded so K steps end within minutes
    for _ in range(min(args.warmup, 1)):
        cpu_port_images_per_sec(args.size, sample, W)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_images_per_sec(args.size, sample, W)
    dt = time.perf_counter() - t0
    v = args.steps * sample / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'yolov4 forward+decode+nms {args.size}x{args.size}, {sample} image/step (bounded sample of batch {args.batch})'},
            'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{sample} image per step x {args.steps} steps, numpy+OpenBLAS oracle (tf.keras not installable)'},
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
    clocks.t0 = time.time()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    l0 = eng.launch_count()
    eng.timer_begin()
    for _ in range(args.steps):
        step_resident()
    ms = eng.timer_end()
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
    clocks.t1 = time.time()
    clk = clocks.stop()
    pk = peaks()
    groups = {}
    layers = eng.layers()
    li = 0
    for t in prof:
        if li == 75 and len(prof) == len(layers) + 1 and 'spp' not in groups:
            groups['spp'] = [1, float(t), 0.0]
            continue
        l = layers[li]; li += 1
        if l['kernel_kind'] in (1, 2):
            nepi, lean = (4, 1) if l['tc_epi_warps'] == 44 else (l['tc_epi_warps'], 0)
            name = (f"conv_tc2_kernel<{l['tile_n']}, {nepi}>" if l['tc_mode'] == 4
                    else f"conv_tc_kernel<{l['tile_n']}, {l['tc_bk']}, 0, {nepi}, {lean}>")
        else:
            name = {4: 'conv0_tc_kernel', 3: 'conv0_direct_kernel', 0: 'conv_simt_kernel'}.get(l['kernel_kind'], 'other')
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
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
                'traffic': traffic, 'kernel': top, 'launches_per_step': n_top, 'avg_launch_us': 1e3 * ms_top / n_top,
                'algorithmic_flops_per_launch': fl_top / n_top, 'share_of_forward': ms_top / float(prof.sum()),
                'traffic_source': 'profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of this kernel)' if traffic else None,
                'peak_source': pk['source'],
                'conv_stack': {'achieved': gflop_step / fwd_ms, 'frac': gflop_step / fwd_ms / pk['tflops'], 'unit': 'TFLOP/s',
                               'flops_per_step': gflop_step * 1e9, 'forward_ms_per_step': fwd_ms,
                               'what': 'all 110 convs + SPP of one step (graph replay), algorithmic conv FLOPs / CUDA-event time'}}

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
            'decode_nms_us_per_img': 1e3 * dn_ms / B, 'valid_detections_img0': int(out[3][0])}
    if not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        v, dt = cpu_port_images_per_sec(S, 2, W)
        line['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                'sample': f'2 images {S}x{S} through the numpy+OpenBLAS oracle ({dt:.1f} s); tf.keras itself is not installable here'}
    emit(line)


if __name__ == '__main__':
    main()
