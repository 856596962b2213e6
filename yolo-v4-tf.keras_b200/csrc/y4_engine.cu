// liby4.so — engine host side: graph planning, darknet loader, launch schedule, C-ABI (include/y4.h).
//
// Graph planning follows the reference builder order (custom_layers.py:100-198) so conv index ==
// Keras creation order == darknet file order (utils.py:12-53).  Planning decisions (all static):
//   * every Concatenate is eliminated: producers write straight into channel slices of the concat buffer;
//   * every residual Add is fused into the epilogue of the producing 3x3 conv (add AFTER activation);
//   * both UpSampling2D are fused into the producing 1x1 conv's store (2x2 replicated write);
//   * BatchNorm (eps 1e-3) is folded into weights/bias at load time;
//   * SPP is one kernel writing the three pooled slices next to its input inside the concat buffer.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <nccl.h>   // types only: the library is dlopen()ed lazily (see NcclApi) so that liby4.so never pins a
                    // libnccl.so.2 before a host process (e.g. torch) loads its own
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/y4.h"
#include "kernels_simt.cuh"
#include "decode_nms.cuh"
#include "conv_tc.cuh"
#include "conv_tc2.cuh"
#include "conv0_tc.cuh"
#include "stem_tc.cuh"
#include "preprocess.cuh"

using namespace y4;

namespace {

thread_local std::string g_create_error;

struct Buf {
    void* ptr = nullptr;
    void* ptr_lo = nullptr;    // split precision (Y4_PREC_FP16X3): low-order fp16 plane, same shape
    int H = 0, W = 0, C = 0;   // logical spatial dims, total channels (ld)
    int elt = 0;               // bytes per element
    size_t bytes = 0;
};
struct View {
    int buf = -1, choff = 0, C = 0, H = 0, W = 0;
};
struct SymOp {
    int kind = 0;              // 0 conv, 1 add, 2 concat, 3 maxpool, 4 upsample
    std::string out;
    std::vector<std::string> ins;
    int idx = -1, cin = 0, cout = 0, k = 0, stride = 1, bn = 1, act = 0, pool = 0;
    int channels = 0, scale = 1;
};

struct ConvOp {
    int idx = 0, cin = 0, cout = 0, cout_pad = 0, k = 1, stride = 1, bn = 1, act = 0;
    int K = 0;
    int raw_in = 0;
    View in, out, res;
    int has_res = 0, upsample = 0, out_f32 = 0;
    int N_OH = 0;              // output spatial size
    float* d_w32 = nullptr;    // [K][cout_pad]
    float* d_bias = nullptr;   // [cout_pad]
    __half* d_w16 = nullptr;   // [cout_pad][K]  (tcgen05 path)
    __half* d_w16_lo = nullptr; // split precision: fp16(w*s - fp16(w*s)), s = per-cout power of two
    float* d_wscale = nullptr;  // split precision: 1/s per cout (exact)
    __half* d_w16k32 = nullptr; // conv 0 only: [32][32], K zero-padded 27 -> 32 (conv0_tc.cuh)
    __half* d_w16_pair = nullptr; // 3x3 stride-2 conv with cin = 32 (conv 1): [cout_pad][6*64] for the pixel-pair view (TcConvDesc::pairx)
    int chain = -1;                   // >= 0: the 1x1 conv with this index runs inside this conv's kernel, on the output tile (chain fusion)
    int fused_a = -1, fused_b = -1;   // >= 0: this entry is the sibling fusion of convs a and b (same input, one GEMM, cout = 2C):
    View out2;                        // columns [0, C) -> out (= a's destination), [C, 2C) -> out2 (= b's); appended after the 110 real convs
    int kind = 0;              // 0 simt, 1 tc flat, 2 tc box
    TcConvPlan tc;             // tensor maps + tile config (conv_tc.cuh)
    TcConvDesc desc{};         // what tc was planned from (the autotuner re-plans candidates from it)
    std::string out_name;
};
struct Step { int type; int conv; };   // type 0 conv (or fused sibling pair), 1 spp, 3 stem (conv 0 + conv 1 in one kernel)

}  // namespace

struct y4_engine {
    y4_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    int elt = 2;                       // activation element size
    std::vector<Buf> bufs;
    std::map<std::string, View> views;
    std::vector<ConvOp> convs;
    std::vector<Step> steps;
    StemPlan stem;                     // conv 0 + conv 1 fused (stem_tc.cuh), when eligible
    View spp_view;                     // concat buffer view of the SPP
    int spp_C = 0;
    float* d_img = nullptr;
    int g[3] = {0, 0, 0};
    int N = 0;                         // boxes per image
    int head_ld = 0;
    int head_buf[3] = {-1, -1, -1};
    float* d_user_heads[3] = {nullptr, nullptr, nullptr};
    float* d_obj[3] = {nullptr, nullptr, nullptr};     // compact objectness logits [3][rows] per scale, written by the tcgen05 head convs
    long long obj_rows[3] = {0, 0, 0};
    bool obj_valid = false;                            // all three head convs run a tcgen05 plan that writes d_obj
    unsigned long long* d_cand_keys = nullptr;
    unsigned long long* d_win_keys = nullptr;          // NMS workspace of the overflow path (decode_nms.cuh)
    int* d_nwin = nullptr;
    unsigned long long* d_part_keys = nullptr;         // nms_image_kernel: per-part survivor lists, [B][kMaxParts][kMaxBoxesCap]
    int* d_ticket = nullptr;                           // [B], = d_cand_count + B: cleared with it at the start of every decode
    int* d_cand_count = nullptr;
    float4* d_boxes = nullptr;
    float* d_out_boxes = nullptr; float* d_out_scores = nullptr; float* d_out_classes = nullptr;
    int* d_out_valid = nullptr; int* d_out_idx = nullptr;
    bool weights_loaded = false;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float4* d_flush = nullptr; size_t flush_elems = 0;
    ncclComm_t comm = nullptr; int rank = 0, nranks = 1;
    void* d_gather = nullptr; size_t gather_bytes = 0;
    std::map<int, std::pair<cudaGraphExec_t, int64_t>> graphs;   // key = (batch*4 + what)*2 + input slot -> (exec, kernels per replay)
    // pipelined host path (y4_submit / y4_collect): two input buffers, H2D on its own stream, results staged in pinned memory
    float* d_img_slot[2] = {nullptr, nullptr};
    int img_slot = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    char* stage[2] = {nullptr, nullptr};          // pinned: boxes | scores | classes | idx | valid
    int64_t n_submitted = 0, n_collected = 0;
    int sub_batch[2] = {0, 0};
    // raw 8-bit input path (y4_predict_u8 / y4_submit_u8): device staging for the source images, one per input slot
    uint8_t* d_u8[2] = {nullptr, nullptr}; size_t u8_cap[2] = {0, 0};
    PreImage* d_pre[2] = {nullptr, nullptr}; PreImage* h_pre[2] = {nullptr, nullptr};
    bool div255_ready = false;
};

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // an already-loaded libnccl.so.2 (torch's) is reused
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd && api.GetErrorString;
        }
    }
    return api;
}

int fail(y4_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_create_error = msg;
    return code;
}
#define CUDA_TRY(e, call)                                                                         \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return fail((e), Y4_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
    } while (0)

// ---- symbolic builder (mirror of custom_layers.py:5-198; kept as data, cf. oracle/netspec.py) -------
struct Builder {
    std::vector<SymOp> ops;
    std::map<std::string, std::pair<int, int>> meta;   // name -> (channels, scale)
    int nconv = 0, nadd = 0, ncat = 0;
    Builder() { meta["img"] = {3, 1}; }
    std::string conv(const std::string& x, int filters, int k, bool down = false, int act = ACT_LEAKY, bool bn = true) {
        SymOp o; o.kind = 0; o.out = "c" + std::to_string(nconv); o.ins = {x};
        o.idx = nconv++; o.cin = meta[x].first; o.cout = filters; o.k = k; o.stride = down ? 2 : 1;
        o.bn = bn; o.act = act; o.channels = filters; o.scale = meta[x].second * (down ? 2 : 1);
        meta[o.out] = {o.channels, o.scale}; ops.push_back(o); return o.out;
    }
    std::string add(const std::string& a, const std::string& b) {
        SymOp o; o.kind = 1; o.out = "r" + std::to_string(++nadd); o.ins = {a, b};
        o.channels = meta[a].first; o.scale = meta[a].second;
        meta[o.out] = {o.channels, o.scale}; ops.push_back(o); return o.out;
    }
    std::string concat(const std::vector<std::string>& parts) {
        SymOp o; o.kind = 2; o.out = "cat" + std::to_string(++ncat); o.ins = parts;
        o.scale = meta[parts[0]].second; o.channels = 0;
        for (auto& p : parts) o.channels += meta[p].first;
        meta[o.out] = {o.channels, o.scale}; ops.push_back(o); return o.out;
    }
    std::string maxpool(const std::string& x, int size) {
        SymOp o; o.kind = 3; o.out = "mp" + std::to_string(size); o.ins = {x}; o.pool = size;
        o.channels = meta[x].first; o.scale = meta[x].second;
        meta[o.out] = {o.channels, o.scale}; ops.push_back(o); return o.out;
    }
    std::string upsample(const std::string& x) {
        SymOp o; o.kind = 4; o.out = "up_" + x; o.ins = {x};
        o.channels = meta[x].first; o.scale = meta[x].second / 2;
        meta[o.out] = {o.channels, o.scale}; ops.push_back(o); return o.out;
    }
    std::string residual(const std::string& x, int f1, int f2, int act) {          // custom_layers.py:34-44
        std::string y = conv(x, f1, 1, false, act);
        y = conv(y, f2, 3, false, act);
        return add(x, y);
    }
    std::string csp(std::string x, int out, int repeat, bool bottleneck = false) {  // custom_layers.py:47-69
        std::string route = conv(x, out, 1, false, ACT_MISH);
        x = conv(x, out, 1, false, ACT_MISH);
        for (int i = 0; i < repeat; i++) x = residual(x, bottleneck ? out / 2 : out, out, ACT_MISH);
        x = conv(x, out, 1, false, ACT_MISH);
        return concat({x, route});
    }
    void build(int num_classes, std::string heads[3]) {
        std::string x = conv("img", 32, 3);                       // custom_layers.py:101 (leaky by default arg)
        x = conv(x, 64, 3, true);
        x = csp(x, 64, 1, true);
        x = conv(x, 64, 1, false, ACT_MISH);
        x = conv(x, 128, 3, true, ACT_MISH);
        x = csp(x, 64, 2);
        x = conv(x, 128, 1, false, ACT_MISH);
        x = conv(x, 256, 3, true, ACT_MISH);
        x = csp(x, 128, 8);
        x = conv(x, 256, 1, false, ACT_MISH);
        std::string route0 = x;
        x = conv(x, 512, 3, true, ACT_MISH);
        x = csp(x, 256, 8);
        x = conv(x, 512, 1, false, ACT_MISH);
        std::string route1 = x;
        x = conv(x, 1024, 3, true, ACT_MISH);
        x = csp(x, 512, 4);
        x = conv(x, 1024, 1, false, ACT_MISH);
        x = conv(x, 512, 1); x = conv(x, 1024, 3); x = conv(x, 512, 1);
        std::string m13 = maxpool(x, 13), m9 = maxpool(x, 9), m5 = maxpool(x, 5);
        x = concat({m13, m9, m5, x});                             // custom_layers.py:130-134
        x = conv(x, 512, 1); x = conv(x, 1024, 3);
        std::string route2 = conv(x, 512, 1);
        const int nout = 3 * (num_classes + 5);
        // neck, custom_layers.py:141-198
        x = conv(route2, 256, 1);
        std::string up = upsample(x);
        std::string r1 = conv(route1, 256, 1);
        x = concat({r1, up});
        const int seq1[5][2] = {{256, 1}, {512, 3}, {256, 1}, {512, 3}, {256, 1}};
        for (auto& s : seq1) x = conv(x, s[0], s[1]);
        std::string route1b = x;
        x = conv(x, 128, 1);
        up = upsample(x);
        std::string r0 = conv(route0, 128, 1);
        x = concat({r0, up});
        const int seq0[5][2] = {{128, 1}, {256, 3}, {128, 1}, {256, 3}, {128, 1}};
        for (auto& s : seq0) x = conv(x, s[0], s[1]);
        std::string route0b = x;
        x = conv(x, 256, 3);
        heads[0] = conv(x, nout, 1, false, ACT_LINEAR, false);
        x = conv(route0b, 256, 3, true);
        x = concat({x, route1b});
        for (auto& s : seq1) x = conv(x, s[0], s[1]);
        std::string route1c = x;
        x = conv(x, 512, 3);
        heads[1] = conv(x, nout, 1, false, ACT_LINEAR, false);
        x = conv(route1c, 512, 3, true);
        x = concat({x, route2});
        const int seq2[5][2] = {{512, 1}, {1024, 3}, {512, 1}, {1024, 3}, {512, 1}};
        for (auto& s : seq2) x = conv(x, s[0], s[1]);
        x = conv(x, 1024, 3);
        heads[2] = conv(x, nout, 1, false, ACT_LINEAR, false);
    }
};

int alloc_buf(y4_engine* e, int H, int W, int C, int elt) {
    Buf b; b.H = H; b.W = W; b.C = C; b.elt = elt;
    b.bytes = (size_t)e->cfg.max_batch * (H + 2) * (W + 2) * C * elt;
    if (cudaMalloc(&b.ptr, b.bytes) != cudaSuccess) return -1;
    if (cudaMemset(b.ptr, 0, b.bytes) != cudaSuccess) return -1;     // halo stays zero forever
    if (e->cfg.precision == Y4_PREC_FP16X3 && elt == 2) {
        if (cudaMalloc(&b.ptr_lo, b.bytes) != cudaSuccess) return -1;
        if (cudaMemset(b.ptr_lo, 0, b.bytes) != cudaSuccess) return -1;
    }
    e->bufs.push_back(b);
    return (int)e->bufs.size() - 1;
}

int plan_graph(y4_engine* e) {
    const int S = e->cfg.img_size, nc = e->cfg.num_classes;
    Builder b; std::string heads[3];
    b.build(nc, heads);
    std::map<std::string, View> placed;         // tensors that live inside a concat buffer
    std::map<std::string, std::string> fused_add, res_of, up_src;
    for (auto& o : b.ops) {
        if (o.kind == 2) {
            int hw = S / o.scale;
            int bi = alloc_buf(e, hw, hw, o.channels, e->elt);
            if (bi < 0) return fail(e, Y4_ERR_CUDA, "cudaMalloc failed for concat buffer " + o.out);
            View v; v.buf = bi; v.choff = 0; v.C = o.channels; v.H = v.W = hw;
            e->views[o.out] = v;
            int off = 0;
            for (auto& part : o.ins) {
                View pv = v; pv.choff = off; pv.C = b.meta[part].first;
                placed[part] = pv; off += pv.C;
            }
        }
    }
    for (auto& o : b.ops) {
        if (o.kind == 1) { fused_add[o.ins[1]] = o.out; res_of[o.ins[1]] = o.ins[0]; }
        if (o.kind == 4) { up_src[o.ins[0]] = o.out; }
    }
    View img; img.buf = -1; img.C = 3; img.H = img.W = S;
    e->views["img"] = img;
    bool spp_done = false;
    for (auto& o : b.ops) {
        if (o.kind == 0) {
            ConvOp c;
            c.idx = o.idx; c.cin = o.cin; c.cout = o.cout; c.k = o.k; c.stride = o.stride; c.bn = o.bn; c.act = o.act;
            c.K = o.k * o.k * o.cin; c.cout_pad = (o.cout + 63) / 64 * 64;
            c.raw_in = (o.ins[0] == "img");
            if (!e->views.count(o.ins[0])) return fail(e, Y4_ERR_ARG, "planner: input not materialised: " + o.ins[0]);
            c.in = e->views[o.ins[0]];
            const int hw = S / o.scale;
            c.N_OH = hw;
            c.out_name = o.out;
            if (fused_add.count(o.out)) {
                int bi = alloc_buf(e, hw, hw, o.cout, e->elt);
                if (bi < 0) return fail(e, Y4_ERR_CUDA, "cudaMalloc failed");
                View v; v.buf = bi; v.C = o.cout; v.H = v.W = hw;
                c.out = v; c.has_res = 1; c.res = e->views[res_of[o.out]];
                c.out_name = fused_add[o.out];
                e->views[c.out_name] = v;
            } else if (up_src.count(o.out)) {
                const std::string& upn = up_src[o.out];
                if (!placed.count(upn)) return fail(e, Y4_ERR_ARG, "planner: upsample not feeding a concat");
                c.out = placed[upn]; c.upsample = 1; c.out_name = upn;
                e->views[upn] = c.out;
            } else if (placed.count(o.out)) {
                c.out = placed[o.out];
                e->views[o.out] = c.out;
            } else {
                const bool head = !o.bn;
                int bi = alloc_buf(e, hw, hw, head ? c.cout_pad : o.cout, head ? 4 : e->elt);
                if (bi < 0) return fail(e, Y4_ERR_CUDA, "cudaMalloc failed");
                View v; v.buf = bi; v.C = o.cout; v.H = v.W = hw;
                c.out = v; c.out_f32 = head;
                e->views[o.out] = v;
            }
            e->convs.push_back(c);
            e->steps.push_back({0, (int)e->convs.size() - 1});
        } else if (o.kind == 3) {
            if (!spp_done) {
                spp_done = true;
                View x = e->views[o.ins[0]];
                if (!placed.count("mp13") || placed["mp13"].choff != 0 || placed["mp9"].choff != x.C ||
                    placed["mp5"].choff != 2 * x.C || x.choff != 3 * x.C || placed["mp13"].buf != x.buf)
                    return fail(e, Y4_ERR_ARG, "planner: unexpected SPP layout");
                e->spp_view = x; e->spp_C = x.C;
                e->steps.push_back({1, -1});
            }
            e->views[o.out] = placed[o.out];
        }
    }
    // Sibling fusion (tcgen05 fp16 mode): csp_block's route and main 1x1 convs read the same tensor (custom_layers.py:59-60:
    // c2|c3, c9|c10, c18|c19, c39|c40, c60|c61).  They run as ONE GEMM with the two weight matrices stacked along N: the input is
    // read once and the N tile doubles; the epilogue stores the two halves to their own destinations.  Same K order and the same
    // per-element arithmetic as the separate convs, so the outputs are bit-identical (tests/test_gpu_determinism.py).
    const bool fuse = e->cfg.precision == Y4_PREC_FP16 && !(getenv("Y4_SIBLING") && getenv("Y4_SIBLING")[0] == '0');
    if (fuse) {
        const size_t nreal = e->convs.size();
        for (size_t i = 0; i + 1 < nreal; i++) {
            const ConvOp a = e->convs[i], b = e->convs[i + 1];
            const bool same_in = a.in.buf == b.in.buf && a.in.choff == b.in.choff && a.in.C == b.in.C && !a.raw_in;
            if (!same_in || a.k != 1 || b.k != 1 || a.stride != 1 || b.stride != 1 || a.cout != b.cout || a.act != b.act || !a.bn || !b.bn ||
                a.has_res || b.has_res || a.upsample || b.upsample || a.out_f32 || b.out_f32 || a.cout % 64 != 0) continue;
            ConvOp f = a;
            f.idx = (int)e->convs.size(); f.cout = 2 * a.cout; f.cout_pad = 2 * a.cout; f.out2 = b.out;
            f.fused_a = (int)i; f.fused_b = (int)i + 1;
            f.out_name = a.out_name + "+" + b.out_name;
            e->convs.push_back(f);
            const int fi = (int)e->convs.size() - 1;
            for (size_t k = 0; k < e->steps.size(); k++)
                if (e->steps[k].type == 0 && e->steps[k].conv == (int)i) e->steps[k].conv = fi;
            for (size_t k = 0; k < e->steps.size(); k++)
                if (e->steps[k].type == 0 && e->steps[k].conv == (int)i + 1) { e->steps.erase(e->steps.begin() + k); break; }
            i++;
        }
    }
    for (int i = 0; i < 3; i++) {
        e->head_buf[i] = e->views[heads[i]].buf;
        e->g[i] = S / e->cfg.strides[i];
        if (e->views[heads[i]].H != e->g[i]) return fail(e, Y4_ERR_ARG, "strides do not match the network's head scales (8,16,32)");
    }
    e->head_ld = e->bufs[e->head_buf[0]].C;
    e->N = 3 * (e->g[0] * e->g[0] + e->g[1] * e->g[1] + e->g[2] * e->g[2]);
    return Y4_OK;
}

template <typename TIn, typename TOut>
void launch_simt_t(y4_engine* e, const ConvOp& c, const SimtConvParams& p) {
    dim3 grid((unsigned)((p.M + 63) / 64), (unsigned)((c.cout + 63) / 64));
    if (c.raw_in) conv_simt_kernel<float, TOut, true><<<grid, 256, 0, e->stream>>>(p);
    else conv_simt_kernel<TIn, TOut, false><<<grid, 256, 0, e->stream>>>(p);
}

void launch_simt(y4_engine* e, const ConvOp& c, int batch) {
    SimtConvParams p{};
    if (c.raw_in) {
        p.in = e->d_img; p.in_ld = 3; p.in_choff = 0; p.in_Hp = e->cfg.img_size; p.in_Wp = e->cfg.img_size;
    } else {
        const Buf& ib = e->bufs[c.in.buf];
        p.in = ib.ptr; p.in_ld = ib.C; p.in_choff = c.in.choff; p.in_Hp = ib.H + 2; p.in_Wp = ib.W + 2;
    }
    p.w = c.d_w32; p.bias = c.d_bias; p.K = c.K; p.cin = c.cin; p.ksize = c.k; p.stride = c.stride;
    p.cout = c.cout; p.cout_pad = c.cout_pad;
    const Buf& ob = e->bufs[c.out.buf];
    p.out = ob.ptr; p.out_ld = ob.C; p.out_choff = c.out.choff;
    p.N = batch; p.OH = c.N_OH; p.OW = c.N_OH;
    if (c.has_res) { const Buf& rb = e->bufs[c.res.buf]; p.res = rb.ptr; p.res_ld = rb.C; p.res_choff = c.res.choff; }
    p.act = c.act; p.upsample = c.upsample; p.out_f32 = c.out_f32;
    p.M = (long long)batch * (c.N_OH + 2) * (c.N_OH + 2);
    if (e->elt == 4) launch_simt_t<float, float>(e, c, p);
    else launch_simt_t<__half, __half>(e, c, p);
    e->launches++;
}

void launch_spp(y4_engine* e, int batch) {
    const Buf& b = e->bufs[e->spp_view.buf];
    SppParams p{b.ptr, b.C, batch, b.H, b.W, e->spp_C};
    long long total = (long long)batch * b.H * b.W * e->spp_C;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (e->elt == 4) spp_kernel<float><<<blocks, 256, 0, e->stream>>>(p);
    else if (e->cfg.precision == Y4_PREC_FP16X3) spp_split_kernel<<<blocks, 256, 0, e->stream>>>(p, b.ptr_lo);
    else if (e->cfg.precision == Y4_PREC_FP16 && e->spp_C % kSppChunk == 0 && (size_t)b.H * b.W * 256 <= 200 * 1024) {
        const size_t smem = (size_t)b.H * b.W * 256;                    // x + three row-max tiles, 64 B per position each
        static DeviceOnce once;                                        // the attribute is per device
        if (once.first_use()) cudaFuncSetAttribute(spp_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        spp_sep_kernel<<<(unsigned)(batch * (e->spp_C / kSppChunk)), 256, smem, e->stream>>>(p);
    } else if (e->cfg.precision == Y4_PREC_FP16 && e->spp_C % 8 == 0)
        spp_half8_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, e->stream>>>(p);
    else spp_kernel<__half><<<blocks, 256, 0, e->stream>>>(p);
    e->launches++;
}

void launch_conv0_direct(y4_engine* e, const ConvOp& c, int batch) {
    const int S = e->cfg.img_size;
    Conv0Params p{e->d_img, c.d_w32, c.d_bias, (__half*)e->bufs[c.out.buf].ptr, (__half*)e->bufs[c.out.buf].ptr_lo, batch, S, c.cout_pad};
    dim3 grid((unsigned)((S + 31) / 32), (unsigned)((S + 15) / 16), (unsigned)batch);
    conv0_direct_kernel<<<grid, 256, 0, e->stream>>>(p);
    e->launches++;
}

void launch_conv0_tc(y4_engine* e, const ConvOp& c, int batch) {
    const int S = e->cfg.img_size;
    Conv0TcParams p{e->d_img, c.d_w16k32, c.d_bias, (__half*)e->bufs[c.out.buf].ptr, batch, S, (S + 127) / 128, 0};
    p.num_tiles = batch * S * p.tiles_per_row;
    const int max_ctas = sm_count() * 12;
    conv0_tc_kernel<<<p.num_tiles < max_ctas ? p.num_tiles : max_ctas, kC0Threads, 0, e->stream>>>(p);
    e->launches++;
}

int run_conv(y4_engine* e, const ConvOp& c, int batch) {
    if (c.kind == 3) { launch_conv0_direct(e, c, batch); return Y4_OK; }
    if (c.kind == 4) { launch_conv0_tc(e, c, batch); return Y4_OK; }
    if (c.kind == 0) { launch_simt(e, c, batch); return Y4_OK; }
    int rc = c.tc.cta2 ? tc2_launch(c.tc, batch, e->stream) : tc_launch(c.tc, batch, e->stream);
    if (rc != 0) return fail(e, Y4_ERR_CUDA, "tcgen05 conv launch failed for conv " + std::to_string(c.idx));
    e->launches++;
    return Y4_OK;
}

int run_step(y4_engine* e, const Step& s, int batch) {
    if (s.type == 0) return run_conv(e, e->convs[s.conv], batch);
    if (s.type == 3) {
        if (stem_launch(e->stem, e->d_img, batch, e->stream) != 0) return fail(e, Y4_ERR_CUDA, "stem (conv 0 + conv 1) launch failed");
        e->launches++;
        return Y4_OK;
    }
    launch_spp(e, batch);
    return Y4_OK;
}

int run_forward(y4_engine* e, int batch) {
    for (auto& s : e->steps) { int rc = run_step(e, s, batch); if (rc) return rc; }
    CUDA_TRY(e, cudaGetLastError());
    return Y4_OK;
}

int run_decode_nms(y4_engine* e, int batch, float iou_thr, float score_thr, const float* const* user_heads) {
    DecodeParams d{};
    const int nc = e->cfg.num_classes;
    int cells = 0, boxes = 0;
    for (int i = 0; i < 3; i++) {
        if (user_heads) { d.head[i] = user_heads[i]; d.ld[i] = 3 * (5 + nc); d.padded[i] = 0; }
        else { d.head[i] = (const float*)e->bufs[e->head_buf[i]].ptr; d.ld[i] = e->head_ld; d.padded[i] = 1; }
        d.obj[i] = (!user_heads && e->obj_valid) ? e->d_obj[i] : nullptr; d.obj_rows[i] = e->obj_rows[i];
        d.g[i] = e->g[i]; d.stride[i] = (float)e->cfg.strides[i];
        d.xyscale[i] = (float)e->cfg.xyscale[i];
        d.xyoff[i] = (float)(0.5 * (e->cfg.xyscale[i] - 1.0));
        d.cell_off[i] = cells; d.box_off[i] = boxes;
        cells += e->g[i] * e->g[i]; boxes += 3 * e->g[i] * e->g[i];
    }
    d.cell_off[3] = cells;
    for (int i = 0; i < 18; i++) d.anchors[i] = e->cfg.anchors[i];
    d.nc = nc; d.C = 5 + nc; d.N = e->N; d.batch = batch; d.img_size = (float)e->cfg.img_size; d.score_thr = score_thr;
    {   // logits below this cannot pass the threshold (DecodeParams::logit_lo)
        const double t = (double)score_thr;
        if (!(t > 0.0)) d.logit_lo = -INFINITY;
        else if (!(t < 1.0)) d.logit_lo = INFINITY;
        else { const double lt = log(t / (1.0 - t)); d.logit_lo = (float)(lt - 1e-3 - 1e-3 * fabs(lt)); }
    }
    d.cand_keys = e->d_cand_keys; d.cand_count = e->d_cand_count; d.boxes = e->d_boxes;
    // candidate counts AND the NMS parts' tickets (contiguous): a launch that died half way must not leave a ticket behind
    CUDA_TRY(e, cudaMemsetAsync(e->d_cand_count, 0, sizeof(int) * 2 * e->cfg.max_batch, e->stream));
    decode_filter_kernel<<<dim3((unsigned)((e->N + kFilterThreads - 1) / kFilterThreads), (unsigned)batch), kFilterThreads, 0, e->stream>>>(d);
    NmsParams n{};
    n.cand_keys = e->d_cand_keys; n.cand_count = e->d_cand_count; n.boxes = e->d_boxes;
    n.N = e->N; n.nc = nc; n.max_boxes = e->cfg.max_boxes; n.iou_thr = iou_thr;
    n.out_boxes = e->d_out_boxes; n.out_scores = e->d_out_scores; n.out_classes = e->d_out_classes;
    n.out_valid = e->d_out_valid; n.out_idx = e->d_out_idx;
    n.win_keys = e->d_win_keys; n.nwin = e->d_nwin; n.part_keys = e->d_part_keys; n.ticket = e->d_ticket;
    n.merge_batch = batch;
    {
        static DeviceOnce once;                                        // the attribute is per device
        if (once.first_use()) cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kImgSmemBytes);
    }
    int parts = sm_count() / batch;                                    // CTAs per image: all of them resident at once
    parts = parts < 1 ? 1 : (parts > kMaxParts ? kMaxParts : parts);
    if (parts > nc) parts = nc;
    nms_image_kernel<<<dim3((unsigned)parts, (unsigned)batch), kImgThreads, kImgSmemBytes, e->stream>>>(n);
    // images with more than kCandCap candidates (none in the usual case: the CTAs return at once) are redone exactly, from the heads
    nms_overflow_kernel<<<batch * nc, kOverflowThreads, sizeof(unsigned) * ((e->N + 31) / 32), e->stream>>>(d, n);
    nms_merge_kernel<<<(batch + kMergeThreads / 32 - 1) / (kMergeThreads / 32), kMergeThreads, 0, e->stream>>>(n);
    e->launches += 4;
    CUDA_TRY(e, cudaGetLastError());
    return Y4_OK;
}

int fetch(y4_engine* e, int batch, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx) {
    const int mb = e->cfg.max_boxes;
    if (boxes) CUDA_TRY(e, cudaMemcpyAsync(boxes, e->d_out_boxes, sizeof(float) * 4 * mb * batch, cudaMemcpyDeviceToHost, e->stream));
    if (scores) CUDA_TRY(e, cudaMemcpyAsync(scores, e->d_out_scores, sizeof(float) * mb * batch, cudaMemcpyDeviceToHost, e->stream));
    if (classes) CUDA_TRY(e, cudaMemcpyAsync(classes, e->d_out_classes, sizeof(float) * mb * batch, cudaMemcpyDeviceToHost, e->stream));
    if (valid) CUDA_TRY(e, cudaMemcpyAsync(valid, e->d_out_valid, sizeof(int) * batch, cudaMemcpyDeviceToHost, e->stream));
    if (cand_idx) CUDA_TRY(e, cudaMemcpyAsync(cand_idx, e->d_out_idx, sizeof(int) * mb * batch, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return Y4_OK;
}

int check_batch(y4_engine* e, int batch) {
    if (!e) return Y4_ERR_ARG;
    if (batch < 1 || batch > e->cfg.max_batch) return fail(e, Y4_ERR_ARG, "batch must be in [1, max_batch]");
    return Y4_OK;
}

int upload_weights(y4_engine* e, const unsigned char* data, size_t nbytes) {
    // utils.py:12-53: 5 x int32 header, then per conv [beta,gamma,mean,var | bias] + OIHW weights
    size_t need = 20;
    for (auto& c : e->convs) if (c.fused_a < 0) need += 4ull * ((c.bn ? 4 * c.cout : c.cout) + (size_t)c.cout * c.cin * c.k * c.k);
    if (nbytes != need)
        return fail(e, Y4_ERR_WEIGHTS, "darknet weights: expected " + std::to_string(need) + " bytes, got " + std::to_string(nbytes));
    size_t off = 20;
    std::vector<float> w32, bias;
    std::vector<__half> w16;
    for (auto& c : e->convs) {
        if (c.fused_a >= 0) continue;                       // built from its two halves below
        const float* f = reinterpret_cast<const float*>(data + off);
        std::vector<double> scale(c.cout, 1.0);
        bias.assign(c.cout_pad, 0.f);
        if (c.bn) {
            const float *beta = f, *gamma = f + c.cout, *mean = f + 2 * c.cout, *var = f + 3 * c.cout;
            for (int o = 0; o < c.cout; o++) {
                scale[o] = (double)gamma[o] / std::sqrt((double)var[o] + 1e-3);        // Keras BN eps (custom_layers.py:26)
                bias[o] = (float)((double)beta[o] - (double)mean[o] * scale[o]);
            }
            f += 4 * c.cout; off += 16ull * c.cout;
        } else {
            for (int o = 0; o < c.cout; o++) bias[o] = f[o];
            f += c.cout; off += 4ull * c.cout;
        }
        // file: [cout][cin][kh][kw]  ->  K index (kh*k+kw)*cin + c   (HWIO flattened, utils.py:42)
        w32.assign((size_t)c.K * c.cout_pad, 0.f);
        w16.assign((size_t)c.cout_pad * c.K, __float2half(0.f));
        const int kk = c.k * c.k;
        for (int o = 0; o < c.cout; o++)
            for (int ci = 0; ci < c.cin; ci++)
                for (int t = 0; t < kk; t++) {
                    float v = (float)((double)f[((size_t)o * c.cin + ci) * kk + t] * scale[o]);
                    size_t kidx = (size_t)t * c.cin + ci;
                    w32[kidx * c.cout_pad + o] = v;
                    w16[(size_t)o * c.K + kidx] = __float2half_rn(v);
                }
        off += 4ull * c.cout * c.cin * kk;
        if (e->cfg.precision == Y4_PREC_FP16X3) {
            // hi/lo split of the folded weights, scaled per cout by a power of two so that hi is O(1) and lo stays a
            // normal fp16 number; the epilogue multiplies the accumulator by the (exact) inverse scale
            std::vector<__half> wlo((size_t)c.cout_pad * c.K, __float2half(0.f));
            std::vector<float> inv(c.cout_pad, 1.0f);
            for (int o = 0; o < c.cout; o++) {
                float mx = 0.f;
                for (int k = 0; k < c.K; k++) mx = std::max(mx, std::fabs(w32[(size_t)k * c.cout_pad + o]));
                int ex = 0;
                if (mx > 0.f) std::frexp(mx, &ex);            // mx = m * 2^ex, m in [0.5, 1)
                // weights scaled into [1024, 2048): lo = w*2^-11 stays a NORMAL fp16 for all but |w| < max * 2^-13;
                // activations are stored * 2^8 for the same reason (conv_tc.cuh act_scale); conv 0 reads the raw image
                const float sc = std::ldexp(1.0f, -ex + 1 + 10);
                inv[o] = std::ldexp(1.0f, ex - 1 - 10) * (c.raw_in ? 1.0f : 1.0f / 256.0f);
                for (int k = 0; k < c.K; k++) {
                    const float ws = w32[(size_t)k * c.cout_pad + o] * sc;
                    const __half hi = __float2half_rn(ws);
                    w16[(size_t)o * c.K + k] = hi;
                    wlo[(size_t)o * c.K + k] = __float2half_rn(ws - __half2float(hi));
                }
            }
            CUDA_TRY(e, cudaMemcpy(c.d_w16_lo, wlo.data(), wlo.size() * 2, cudaMemcpyHostToDevice));
            CUDA_TRY(e, cudaMemcpy(c.d_wscale, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
        }
        CUDA_TRY(e, cudaMemcpy(c.d_w32, w32.data(), w32.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(e, cudaMemcpy(c.d_w16, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
        if (c.d_w16_pair) {
            // pixel-pair view: K' = (kh, j, e): j = 0 -> [kw 0 | kw 1], j = 1 -> [kw 2 | zeros], 32 channels each
            std::vector<__half> wp((size_t)c.cout_pad * 384, __float2half(0.f));
            for (int o = 0; o < c.cout; o++)
                for (int kh = 0; kh < 3; kh++)
                    for (int kw = 0; kw < 3; kw++)
                        for (int ci = 0; ci < 32; ci++)
                            wp[(size_t)o * 384 + (kh * 2 + (kw >> 1)) * 64 + (kw & 1) * 32 + ci] = w16[(size_t)o * c.K + (kh * 3 + kw) * 32 + ci];
            CUDA_TRY(e, cudaMemcpy(c.d_w16_pair, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
        }
        if (c.d_w16k32) {
            std::vector<__half> wk(32 * 32, __float2half(0.f));
            for (int o = 0; o < c.cout && o < 32; o++)
                for (int k = 0; k < c.K && k < 32; k++) wk[o * 32 + k] = w16[(size_t)o * c.K + k];
            CUDA_TRY(e, cudaMemcpy(c.d_w16k32, wk.data(), wk.size() * 2, cudaMemcpyHostToDevice));
        }
        CUDA_TRY(e, cudaMemcpy(c.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    }
    for (auto& c : e->convs) {
        if (c.fused_a < 0) continue;                        // sibling fusion: stack the two folded weight matrices / biases along N
        const ConvOp &a = e->convs[c.fused_a], &b = e->convs[c.fused_b];
        const size_t wn = (size_t)a.cout * a.K;
        CUDA_TRY(e, cudaMemcpy(c.d_w16, a.d_w16, wn * 2, cudaMemcpyDeviceToDevice));
        CUDA_TRY(e, cudaMemcpy(c.d_w16 + wn, b.d_w16, wn * 2, cudaMemcpyDeviceToDevice));
        CUDA_TRY(e, cudaMemcpy(c.d_bias, a.d_bias, a.cout * 4, cudaMemcpyDeviceToDevice));
        CUDA_TRY(e, cudaMemcpy(c.d_bias + a.cout, b.d_bias, a.cout * 4, cudaMemcpyDeviceToDevice));
    }
    e->weights_loaded = true;
    return Y4_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int y4_default_config(y4_config* cfg) {
    if (!cfg) return Y4_ERR_ARG;
    memset(cfg, 0, sizeof(*cfg));
    cfg->img_size = 416; cfg->num_classes = 80; cfg->max_batch = 1; cfg->precision = Y4_PREC_FP16; cfg->device = 0;
    cfg->max_boxes = 100;
    const int st[3] = {8, 16, 32};
    const float an[18] = {12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401};
    const double xs[3] = {1.2, 1.1, 1.05};
    for (int i = 0; i < 3; i++) { cfg->strides[i] = st[i]; cfg->xyscale[i] = xs[i]; }
    for (int i = 0; i < 18; i++) cfg->anchors[i] = an[i];
    cfg->iou_threshold = 0.413f; cfg->score_threshold = 0.3f;
    return Y4_OK;
}

const char* y4_last_error(const y4_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int y4_create(y4_engine** out, const y4_config* cfg) {
    if (!out || !cfg) return fail(nullptr, Y4_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->img_size < 32 || cfg->img_size % 32 != 0 || cfg->img_size > 2048)
        return fail(nullptr, Y4_ERR_ARG, "img_size must be a multiple of the last stride (32)  [models.py:24]");
    if (cfg->num_classes < 1 || cfg->num_classes > 255) return fail(nullptr, Y4_ERR_ARG, "num_classes must be in [1,255]");
    if (cfg->max_boxes < 1 || cfg->max_boxes > kMaxBoxesCap || cfg->max_boxes * cfg->num_classes > kSelCap)
        return fail(nullptr, Y4_ERR_ARG, "max_boxes must be in [1,128] and max_boxes*num_classes <= 8192");
    if (cfg->max_batch < 1) return fail(nullptr, Y4_ERR_ARG, "max_batch must be >= 1");
    if (cfg->precision < 0 || cfg->precision > 3) return fail(nullptr, Y4_ERR_ARG, "unknown precision");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, Y4_ERR_CUDA, "no CUDA device available (this engine has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, Y4_ERR_ARG, "bad device ordinal");
    cudaDeviceProp prop;
    if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess)
        return fail(nullptr, Y4_ERR_CUDA, "cudaSetDevice failed");
    if (prop.major != 10)
        return fail(nullptr, Y4_ERR_CUDA, "device is not sm_100 (Blackwell B200); this library ships sm_100a code only");

    y4_engine* e = new y4_engine();
    e->cfg = *cfg;
    e->elt = cfg->precision == Y4_PREC_FP32 ? 4 : 2;
    auto bail = [&](int rc) { g_create_error = e->err; y4_destroy(e); return rc; };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(e, Y4_ERR_CUDA, "stream create failed"));
    cudaEventCreate(&e->ev0); cudaEventCreate(&e->ev1);
    int rc = plan_graph(e);
    if (rc) return bail(rc);
    if (e->N >= (1 << 24)) return bail(fail(e, Y4_ERR_ARG, "too many boxes per image for the 24-bit box field of the NMS keys"));
    const int S = cfg->img_size, B = cfg->max_batch, mb = cfg->max_boxes;
#define CREATE_TRY(call) do { if ((call) != cudaSuccess) return bail(fail(e, Y4_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError()))); } while (0)
    CREATE_TRY(cudaMalloc(&e->d_img, sizeof(float) * 3 * S * S * B));
    e->d_img_slot[0] = e->d_img;
    for (auto& c : e->convs) {
        if (c.fused_a >= 0) {                               // tcgen05 only: fp16 weights + bias
            CREATE_TRY(cudaMalloc(&c.d_w16, sizeof(__half) * c.K * c.cout_pad));
            CREATE_TRY(cudaMalloc(&c.d_bias, sizeof(float) * c.cout_pad));
            continue;
        }
        CREATE_TRY(cudaMalloc(&c.d_w32, sizeof(float) * c.K * c.cout_pad));
        CREATE_TRY(cudaMalloc(&c.d_w16, sizeof(__half) * c.K * c.cout_pad));
        if (!c.raw_in && c.cin == 32 && c.k == 3 && c.stride == 2) CREATE_TRY(cudaMalloc(&c.d_w16_pair, sizeof(__half) * 384 * c.cout_pad));
        if (c.raw_in && c.K == 27 && c.cout == 32) { CREATE_TRY(cudaMalloc(&c.d_w16k32, sizeof(__half) * 32 * 32)); CREATE_TRY(cudaMemset(c.d_w16k32, 0, sizeof(__half) * 32 * 32)); }
        CREATE_TRY(cudaMalloc(&c.d_bias, sizeof(float) * c.cout_pad));
        if (cfg->precision == Y4_PREC_FP16X3) {
            CREATE_TRY(cudaMalloc(&c.d_w16_lo, sizeof(__half) * c.K * c.cout_pad));
            CREATE_TRY(cudaMalloc(&c.d_wscale, sizeof(float) * c.cout_pad));
            CREATE_TRY(cudaMemset(c.d_wscale, 0, sizeof(float) * c.cout_pad));
        }
    }
    for (int i = 0; i < 3; i++) {
        CREATE_TRY(cudaMalloc(&e->d_user_heads[i], sizeof(float) * B * e->g[i] * e->g[i] * 3 * (5 + cfg->num_classes)));
        e->obj_rows[i] = (long long)B * (e->g[i] + 2) * (e->g[i] + 2);
        CREATE_TRY(cudaMalloc(&e->d_obj[i], sizeof(float) * 3 * e->obj_rows[i]));
        CREATE_TRY(cudaMemset(e->d_obj[i], 0, sizeof(float) * 3 * e->obj_rows[i]));
    }
    CREATE_TRY(cudaMalloc(&e->d_cand_keys, sizeof(unsigned long long) * kCandCap * B));
    CREATE_TRY(cudaMalloc(&e->d_cand_count, sizeof(int) * 2 * B));
    e->d_ticket = e->d_cand_count + B;
    CREATE_TRY(cudaMalloc(&e->d_win_keys, sizeof(unsigned long long) * (size_t)cfg->num_classes * mb * B));
    CREATE_TRY(cudaMalloc(&e->d_nwin, sizeof(int) * 256 * B));
    CREATE_TRY(cudaMalloc(&e->d_part_keys, sizeof(unsigned long long) * kMaxParts * kMaxBoxesCap * B));
    CREATE_TRY(cudaMalloc(&e->d_boxes, sizeof(float4) * e->N * B));
    CREATE_TRY(cudaMemset(e->d_boxes, 0, sizeof(float4) * e->N * B));
    CREATE_TRY(cudaMalloc(&e->d_out_boxes, sizeof(float) * 4 * mb * B));
    CREATE_TRY(cudaMalloc(&e->d_out_scores, sizeof(float) * mb * B));
    CREATE_TRY(cudaMalloc(&e->d_out_classes, sizeof(float) * mb * B));
    CREATE_TRY(cudaMalloc(&e->d_out_valid, sizeof(int) * B));
    CREATE_TRY(cudaMalloc(&e->d_out_idx, sizeof(int) * mb * B));
    e->flush_elems = (size_t)(192u << 20) / sizeof(float4);      // 192 MB > 126 MB L2
    CREATE_TRY(cudaMalloc(&e->d_flush, e->flush_elems * sizeof(float4)));
    // tcgen05 plans (tensor maps need the buffer addresses, which are now fixed)
    if (cfg->precision == Y4_PREC_FP16 || cfg->precision == Y4_PREC_FP16X3) {
        const bool split = cfg->precision == Y4_PREC_FP16X3;
        for (auto& c : e->convs) {
            TcConvDesc d{};
            d.cin = c.cin; d.cout = c.cout; d.cout_pad = c.cout_pad; d.k = c.k; d.stride = c.stride; d.act = c.act;
            d.raw_in = c.raw_in; d.max_batch = B; d.OH = c.N_OH;
            if (!c.raw_in) { const Buf& ib = e->bufs[c.in.buf]; d.in = ib.ptr; d.in_ld = ib.C; d.in_choff = c.in.choff; d.in_H = ib.H; }
            const Buf& ob = e->bufs[c.out.buf];
            d.out = ob.ptr; d.out_ld = ob.C; d.out_choff = c.out.choff; d.out_f32 = c.out_f32; d.upsample = c.upsample;
            if (c.has_res) { const Buf& rb = e->bufs[c.res.buf]; d.res = rb.ptr; d.res_ld = rb.C; d.res_choff = c.res.choff; }
            d.w16 = c.d_w16; d.bias = c.d_bias; d.w16_pair = split ? nullptr : c.d_w16_pair;
            if (c.out_f32)
                for (int hi = 0; hi < 3; hi++)
                    if (c.out.buf == e->head_buf[hi]) { d.obj_out = e->d_obj[hi]; d.obj_c0 = 4; d.obj_stride = 5 + cfg->num_classes; d.obj_rows = e->obj_rows[hi]; }
            if (c.fused_a >= 0) { const Buf& ob2 = e->bufs[c.out2.buf]; d.out2 = ob2.ptr; d.out2_ld = ob2.C; d.out2_choff = c.out2.choff; d.split_col = c.cout / 2; }
            if (split) {
                d.split = 1; d.w16_lo = c.d_w16_lo; d.wscale = c.d_wscale; d.out_lo = ob.ptr_lo;
                if (!c.raw_in) d.in_lo = e->bufs[c.in.buf].ptr_lo;
                if (c.has_res) d.res_lo = e->bufs[c.res.buf].ptr_lo;
            }
            std::string terr;
            if (c.raw_in && c.cin == 3 && c.cout == 32 && c.k == 3 && c.act == ACT_LEAKY && c.out.choff == 0 && ob.C == 32) {
                // conv 0: tensor-core kernel (in-register im2col, kind 4) unless Y4_C0=direct (CUDA-core direct conv, kind 3)
                const char* c0 = getenv("Y4_C0");
                c.kind = (c0 && c0[0] == 'd') || !c.d_w16k32 || split ? 3 : 4;   // split precision: fp32 direct conv writing hi + lo
                continue;
            }
            int kind = c.fused_a >= 0 ? tc_plan(d, &c.tc, &terr, 0, 99, 0, 1, /*epi*/ 1, 4, 0, 32, 0) : tc_plan(d, &c.tc, &terr);
            if (c.fused_a >= 0 && kind != 1) return bail(fail(e, Y4_ERR_CUDA, "tcgen05 plan failed for fused conv " + c.out_name + ": " + terr));
            if (kind < 0) return bail(fail(e, Y4_ERR_CUDA, "tcgen05 plan failed for conv " + std::to_string(c.idx) + ": " + terr));
            c.kind = kind;
            c.desc = d;
        }
        // conv 0 + conv 1 as one kernel (stem_tc.cuh): conv 0's output never reaches HBM.  Y4_STEM=0 keeps the two kernels.
        if (!split && !(getenv("Y4_STEM") && getenv("Y4_STEM")[0] == '0') && e->convs.size() > 1) {
            ConvOp &c0 = e->convs[0], &c1 = e->convs[1];
            const bool shape_ok = c0.kind == 4 && c0.d_w16k32 && c1.cin == 32 && c1.cout == 64 && c1.k == 3 && c1.stride == 2 && c1.act == ACT_LEAKY &&
                                  c0.act == ACT_LEAKY && !c1.has_res && !c1.upsample && !c1.out_f32 && c1.in.buf == c0.out.buf && c1.fused_a < 0;
            if (shape_ok) {
                const Buf& ob = e->bufs[c1.out.buf];
                std::string serr;
                if (stem_plan(&e->stem, c0.d_bias, c0.d_w16k32, c1.d_bias, c1.d_w16, ob.ptr, ob.C, c1.out.choff, S, B, &serr)) {
                    for (size_t k = 0; k < e->steps.size(); k++)
                        if (e->steps[k].type == 0 && e->steps[k].conv == 0) e->steps[k] = Step{3, -1};
                    for (size_t k = 0; k < e->steps.size(); k++)
                        if (e->steps[k].type == 0 && e->steps[k].conv == 1) { e->steps.erase(e->steps.begin() + k); break; }
                }
            }
        }
        // Chain fusion (CTA-pair kernel, conv_tc2.cuh): a 1x1 conv whose input is exactly the tensor the previous launch produces
        // (residual_block's first conv on the block input r_k, custom_layers.py:38-39; the first residual after csp_block's main
        // 1x1) runs on the output tile while it is in shared memory: one launch, one read of r_k less.  Correct (bit-identical,
        // tests/test_gpu_determinism.py) but NOT faster on B200: the fused kernel costs what the two launches cost (3x3 128->128 at 76^2:
        // 0.109 ms vs 0.068 + 0.032) or more (1x1 producers: 0.63 vs 0.44 ms), because it removes HBM traffic while both epilogues (the
        // actual bottleneck of the 1x1 layers) still run, now serialised on one tile's TMEM stage.  Opt-in: Y4_CHAIN=1.
        if (!split && getenv("Y4_CHAIN") && getenv("Y4_CHAIN")[0] == '1') {
            for (size_t k = 0; k + 1 < e->steps.size(); k++) {
                if (e->steps[k].type != 0 || e->steps[k + 1].type != 0) continue;
                const int pi = e->steps[k].conv, qi = e->steps[k + 1].conv;
                ConvOp& P = e->convs[pi];
                const ConvOp& Q = e->convs[qi];
                if (P.kind != 1 || P.stride != 1 || P.out_f32 || P.upsample || P.chain >= 0) continue;
                if (Q.k != 1 || Q.stride != 1 || Q.has_res || Q.upsample || Q.out_f32 || Q.fused_a >= 0 || Q.kind != 1) continue;
                const View& src = P.fused_a >= 0 ? P.out2 : P.out;
                if (Q.in.buf != src.buf || Q.in.choff != src.choff || Q.in.C != src.C) continue;
                TcConvDesc d = P.desc;
                const Buf& qb = e->bufs[Q.out.buf];
                d.q_on = 1; d.q_w16 = Q.d_w16; d.q_bias = Q.d_bias; d.q_cin = Q.cin; d.q_cout = Q.cout; d.q_cout_pad = Q.cout_pad;
                d.q_act = Q.act; d.q_col0 = P.fused_a >= 0 ? P.cout / 2 : 0;
                d.q_out = qb.ptr; d.q_out_ld = qb.C; d.q_out_choff = Q.out.choff;
                TcConvPlan pl;
                std::string cerr2;
                if (tc_plan2(d, &pl, &cerr2, P.cout_pad, 224, 2, 8, 32, 0) != 1) continue;
                P.desc = d; P.tc = pl; P.chain = qi;
                e->steps.erase(e->steps.begin() + k + 1);
            }
        }
        // Plan-time autotune, IN CONTEXT: every candidate configuration is planned for all layers it applies to, the whole
        // forward runs with per-layer CUDA events, and each layer keeps the plan that was fastest where it actually sits
        // (inputs in L2 or not, neighbours' tails) -- timing a layer alone on a hot L2 picked plans that lose in the pipeline.
        // Candidates: N tile x shared-memory budget (ring depth; ~110 KB lets two persistent CTAs share an SM) x k-blocks per
        // barrier x {A-patch reuse, resident weights} x epilogue {per-thread stores, slab + TMA store with 4 or 8 warps}.
        // Every candidate accumulates K in the same order and rounds once, so outputs are bit-identical across them
        // (and therefore across GPUs, whatever each one picks).
        struct Cand { int bn, kb, patch, group, epi, nepi, bres, gw, cta2, lean, pairx; };
        std::vector<Cand> cands;
        // single-CTA A-patch plans: since the warp-uniform issue loops the per-kernel-row patch wins on conv 5 (3x3 32->64 at 304^2:
        // 0.346 -> 0.309 ms, three 130-row loads instead of nine 128-row ones); Y4_PATCH=0 removes them from the candidate list
        const bool allow_patch = !(getenv("Y4_PATCH") && getenv("Y4_PATCH")[0] == '0');
        const bool allow_bres = !(getenv("Y4_BRES") && getenv("Y4_BRES")[0] == '0');
        const int epi_mode = getenv("Y4_EPI") ? atoi(getenv("Y4_EPI")) : 1;     // 0 never, 1 autotune, 2 wherever eligible
        const int nepi_mode = getenv("Y4_NEPI") ? atoi(getenv("Y4_NEPI")) : 0;  // 4 / 8: only that many epilogue warps
        const int gw_mode = getenv("Y4_GW") ? atoi(getenv("Y4_GW")) : 0;        // 32 / 64: only that slab group width
        if (const char* f = getenv("Y4_FORCE")) {                                 // "bn,kb,patch,group,epi,nepi,bres,gw": that plan wherever it applies
            Cand cd{}; if (sscanf(f, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d", &cd.bn, &cd.kb, &cd.patch, &cd.group, &cd.epi, &cd.nepi, &cd.bres, &cd.gw, &cd.cta2, &cd.lean, &cd.pairx) >= 8) cands.push_back(cd);
        } else {
            for (int g : {1, 2, 3, 6})                                         // conv 1 through the pixel-pair view: single CTA (resident W or not) and CTA pair
                for (int kb : {112, 224}) {
                    cands.push_back({64, kb, 0, g, 0, 4, 0, 32, 0, 0, 1});
                    cands.push_back({64, kb, 0, g, 0, 4, 1, 32, 0, 0, 1});
                    if (kb == 224) cands.push_back({64, kb, 0, g, 1, 4, 0, 32, 1, 0, 1});
                }
            static const bool allow_cta2 = !(getenv("Y4_CTA2") && getenv("Y4_CTA2")[0] == '0');
            if (allow_cta2)                                                      // CTA-pair kernel: {N tile, k-blocks per barrier, epilogue warps, group width}
                for (int bn : {64, 128, 256})
                    for (int g : {1, 2, 3})
                        for (auto& ep : {std::pair<int, int>{8, 32}, std::pair<int, int>{4, 64}, std::pair<int, int>{4, 32}, std::pair<int, int>{16, 32}})
                            if (ep.first != 16 || (bn >= 128 && !(getenv("Y4_NEPI16") && getenv("Y4_NEPI16")[0] == '0')))
                                cands.push_back({bn, 224, 0, g, 1, ep.first, 0, ep.second, 1});
            if (allow_cta2 && !(getenv("Y4_PATCH2") && getenv("Y4_PATCH2")[0] == '0'))   // CTA pair + A-patch reuse (3x3 stride 1): cuts the L2 -> smem traffic
                for (int bn : {64, 128, 256})
                    for (auto& ep : {std::pair<int, int>{8, 32}, std::pair<int, int>{4, 64}, std::pair<int, int>{4, 32}, std::pair<int, int>{16, 32}})
                        for (int g : {1, 2}) {
                            if (ep.first == 16 && (bn < 128 || (getenv("Y4_NEPI16") && getenv("Y4_NEPI16")[0] == '0'))) continue;
                            // g = 2: deep B ring (tc_plan2: up to 12 weight stages beside three patches).  Measured: no gain on any layer
                            // (the six-stage ring already covers the L2 latency), so it stays out of the default list: Y4_DEEPB=1
                            if (g == 2 && !(getenv("Y4_DEEPB") && getenv("Y4_DEEPB")[0] == '1')) continue;
                            cands.push_back({bn, 224, 1, g, 1, ep.first, 0, ep.second, 1});
                        }
            for (int bres = 0; bres <= (allow_bres ? 1 : 0); bres++)            // lean 4-warp epilogue, four CTAs per SM (64-wide tiles)
                for (int gw : {32, 64})
                    for (int kb : {56, 75}) cands.push_back({64, kb, 0, 1, 1, 4, bres, gw, 0, 1});
            // {epilogue, epilogue warps, group width}
            const int epis[][3] = {{1, 4, 64}, {1, 4, 32}, {1, 8, 32}, {0, 4, 32}};
            for (int bn : {64, 128, 256})
                for (auto& ep : epis) {
                    const int epi = ep[0], nepi = ep[1], gw = ep[2];
                    if (epi && epi_mode == 0) continue;
                    if (epi && nepi_mode && nepi != nepi_mode) continue;
                    if (epi && gw_mode && gw != gw_mode) continue;
                    for (int bres = 0; bres <= (allow_bres ? 1 : 0); bres++) {
                        for (int g : {1, 2}) cands.push_back({bn, 112, 0, g, epi, nepi, bres, gw});       // two CTAs per SM
                        for (int g : {1, 2, 3, 99}) cands.push_back({bn, 224, 0, g, epi, nepi, bres, gw});
                        if (!bres) cands.push_back({bn, 150, 0, 1, epi, nepi, 0, gw});
                    }
                    if (allow_patch)                                            // 1: whole-halo patches, 3: one patch per kernel row
                        for (int pm : {1, 3})
                            for (int bres = 0; bres <= (allow_bres ? 1 : 0); bres++) {
                                cands.push_back({bn, 224, pm, 1, epi, nepi, bres, gw});
                                cands.push_back({bn, 112, pm, 1, epi, nepi, bres, gw});
                            }
                }
        }
        const char* at = getenv("Y4_AUTOTUNE");
        if (!(at && at[0] == '0')) {
            const size_t nsteps = e->steps.size();
            std::vector<cudaEvent_t> ev(nsteps + 1);
            for (auto& x : ev) cudaEventCreate(&x);
            std::vector<float> best_ms(e->convs.size(), 1e30f);
            std::vector<TcConvPlan> best(e->convs.size()), trial(e->convs.size());
            std::vector<char> has(e->convs.size());
            // the three fastest plans seen per layer, re-timed against each other at the end (two timed passes per candidate are
            // noisy at the 3 % level, and near-ties are common)
            constexpr int kTop = 3;
            std::vector<std::array<std::pair<float, TcConvPlan>, kTop>> top(e->convs.size());
            for (auto& t3 : top) for (auto& t1 : t3) t1.first = 1e30f;
            // one forward with per-step events; each conv runs `use[li]` (or its current best); ms[i] = min over `reps` timed passes
            auto timed_forward = [&](const std::vector<TcConvPlan>& use, std::vector<char>& ok, int reps, std::vector<float>& ms) -> bool {
                ms.assign(nsteps, 1e30f);
                for (int rep = 0; rep <= reps; rep++) {                            // rep 0 warms up (and sets the smem attribute)
                    cudaEventRecord(ev[0], e->stream);
                    for (size_t i = 0; i < nsteps; i++) {
                        auto& st = e->steps[i];
                        if (st.type == 0) {
                            ConvOp& c = e->convs[st.conv];
                            const TcConvPlan keep = c.tc;
                            if (ok[st.conv]) c.tc = use[st.conv]; else if (c.kind == 1 || c.kind == 2) c.tc = best[st.conv];
                            const int rc2 = run_conv(e, c, B);
                            c.tc = keep;
                            if (rc2) { cudaGetLastError(); if (ok[st.conv]) ok[st.conv] = 2; }          // 2: launch refused
                        } else if (run_step(e, st, B)) cudaGetLastError();
                        cudaEventRecord(ev[i + 1], e->stream);
                    }
                    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return false;
                    if (rep == 0) continue;
                    for (size_t i = 0; i < nsteps; i++) { float t = 0.f; cudaEventElapsedTime(&t, ev[i], ev[i + 1]); if (t < ms[i]) ms[i] = t; }
                }
                return true;
            };
            const bool forced = getenv("Y4_FORCE") != nullptr;
            for (int ci = -1; ci < (int)cands.size(); ci++) {                      // -1: the default plans
                bool any = ci < 0;
                for (size_t li = 0; li < e->convs.size(); li++) {
                    ConvOp& c = e->convs[li];
                    has[li] = 0;
                    if (c.kind != 1 && c.kind != 2) continue;
                    if (ci < 0) { trial[li] = c.tc; best[li] = c.tc; has[li] = 1; continue; }
                    const Cand& cd = cands[ci];
                    if (!cd.epi && epi_mode == 2 && best[li].p.epi) continue;
                    std::string er2;
                    TcConvDesc dd = c.desc;
                    dd.pairx = cd.pairx;
                    if (cd.pairx && !dd.w16_pair) continue;
                    if (cd.cta2) { if (tc_plan2(dd, &trial[li], &er2, cd.bn, cd.kb, cd.group, cd.nepi, cd.gw, cd.patch) != c.kind) continue; }
                    else if (tc_plan(dd, &trial[li], &er2, cd.bn, cd.kb, cd.patch, cd.group, cd.epi, cd.nepi, cd.bres, cd.gw, cd.lean) != c.kind) continue;
                    has[li] = 1; any = true;
                }
                if (!any) continue;
                std::vector<float> ms;
                if (!timed_forward(trial, has, 2, ms)) return bail(fail(e, Y4_ERR_CUDA, "autotune forward failed (candidate " + std::to_string(ci) + ")"));
                for (size_t i = 0; i < nsteps; i++) {
                    if (e->steps[i].type != 0) continue;
                    const int li = e->steps[i].conv;
                    if (has[li] != 1) continue;
                    if (ms[i] < best_ms[li] || (forced && ci >= 0)) { best_ms[li] = ms[i]; best[li] = trial[li]; }
                    auto& t3 = top[li];
                    for (int k = 0; k < kTop; k++)
                        if (ms[i] < t3[k].first) { for (int m = kTop - 1; m > k; m--) t3[m] = t3[m - 1]; t3[k] = {ms[i], trial[li]}; break; }
                }
            }
            if (!forced) {
                // confirmation: rank r of every layer runs together, four timed passes each; the layer keeps the fastest of its three
                std::vector<float> conf_ms(e->convs.size(), 1e30f);
                for (int r = 0; r < kTop; r++) {
                    for (size_t li = 0; li < e->convs.size(); li++) {
                        has[li] = ((e->convs[li].kind == 1 || e->convs[li].kind == 2) && top[li][r].first < 1e29f) ? 1 : 0;
                        if (has[li]) trial[li] = top[li][r].second;
                    }
                    std::vector<float> ms;
                    if (!timed_forward(trial, has, 4, ms)) return bail(fail(e, Y4_ERR_CUDA, "autotune confirmation pass failed"));
                    for (size_t i = 0; i < nsteps; i++) {
                        if (e->steps[i].type != 0) continue;
                        const int li = e->steps[i].conv;
                        if (has[li] == 1 && ms[i] < conf_ms[li]) { conf_ms[li] = ms[i]; best[li] = trial[li]; }
                    }
                }
            }
            for (size_t li = 0; li < e->convs.size(); li++) if (e->convs[li].kind == 1 || e->convs[li].kind == 2) e->convs[li].tc = best[li];
            for (auto& x : ev) cudaEventDestroy(x);
        }
        // autotune launches wrote act(bias=0)=0-ish garbage into interiors only; halos were never touched.
    }
    {
        int nh = 0;
        for (auto& c : e->convs) nh += (c.out_f32 && (c.kind == 1 || c.kind == 2)) ? 1 : 0;
        e->obj_valid = nh == 3;
    }
    CREATE_TRY(cudaDeviceSynchronize());
    *out = e;
    return Y4_OK;
}

void y4_destroy(y4_engine* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.second.first);
    if (e->comm) nccl().CommDestroy(e->comm);
    for (auto& b : e->bufs) { cudaFree(b.ptr); cudaFree(b.ptr_lo); }
    for (auto& c : e->convs) { cudaFree(c.d_w32); cudaFree(c.d_w16); cudaFree(c.d_w16k32); cudaFree(c.d_w16_pair); cudaFree(c.d_w16_lo); cudaFree(c.d_wscale); cudaFree(c.d_bias); }
    cudaFree(e->d_img_slot[0]); cudaFree(e->d_img_slot[1]);
    for (int i = 0; i < 2; i++) { if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]); if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]); if (e->stage[i]) cudaFreeHost(e->stage[i]); }
    for (int i = 0; i < 2; i++) { cudaFree(e->d_u8[i]); cudaFree(e->d_pre[i]); if (e->h_pre[i]) cudaFreeHost(e->h_pre[i]); }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int i = 0; i < 3; i++) { cudaFree(e->d_user_heads[i]); cudaFree(e->d_obj[i]); }
    cudaFree(e->d_cand_keys); cudaFree(e->d_cand_count); cudaFree(e->d_boxes);
    cudaFree(e->d_win_keys); cudaFree(e->d_nwin); cudaFree(e->d_part_keys);
    cudaFree(e->d_out_boxes); cudaFree(e->d_out_scores); cudaFree(e->d_out_classes);
    cudaFree(e->d_out_valid); cudaFree(e->d_out_idx);
    cudaFree(e->d_flush); cudaFree(e->d_gather);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int y4_load_darknet_from_memory(y4_engine* e, const void* data, size_t nbytes) {
    if (!e || !data) return e ? fail(e, Y4_ERR_ARG, "null data") : Y4_ERR_ARG;
    cudaSetDevice(e->cfg.device);
    return upload_weights(e, static_cast<const unsigned char*>(data), nbytes);
}

int y4_load_darknet(y4_engine* e, const char* path) {
    if (!e || !path) return e ? fail(e, Y4_ERR_ARG, "null path") : Y4_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(e, Y4_ERR_WEIGHTS, std::string("cannot open ") + path);
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> buf((size_t)sz);
    size_t got = fread(buf.data(), 1, (size_t)sz, f);
    fclose(f);
    if (got != (size_t)sz) return fail(e, Y4_ERR_WEIGHTS, "short read");
    return y4_load_darknet_from_memory(e, buf.data(), buf.size());
}

// Every entry point except y4_submit* / y4_collect uses input slot 0 and the shared result buffers, which belong to the batches
// in flight while submits are outstanding.
static int ready(y4_engine* e, int batch, bool need_weights, bool pipelined = false) {
    int rc = check_batch(e, batch);
    if (rc) return rc;
    if (need_weights && !e->weights_loaded) return fail(e, Y4_ERR_STATE, "weights not loaded (call y4_load_darknet first)");
    if (!pipelined && e->n_submitted != e->n_collected) return fail(e, Y4_ERR_STATE, "y4_submit batches are in flight: y4_collect them first");
    cudaSetDevice(e->cfg.device);
    return Y4_OK;
}

int y4_predict(y4_engine* e, const float* imgs, int32_t batch, float* boxes, float* scores, float* classes,
               int32_t* valid, int32_t* cand_idx) {
    int rc = ready(e, batch, true); if (rc) return rc;
    if (!imgs) return fail(e, Y4_ERR_ARG, "null imgs");
    const size_t n = (size_t)batch * e->cfg.img_size * e->cfg.img_size * 3;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_img, imgs, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    rc = run_forward(e, batch); if (rc) return rc;
    rc = run_decode_nms(e, batch, e->cfg.iou_threshold, e->cfg.score_threshold, nullptr); if (rc) return rc;
    return fetch(e, batch, boxes, scores, classes, valid, cand_idx);
}

int y4_forward_heads(y4_engine* e, const float* imgs, int32_t batch, float* hs, float* hm, float* hl) {
    int rc = ready(e, batch, true); if (rc) return rc;
    if (!imgs) return fail(e, Y4_ERR_ARG, "null imgs");
    const size_t n = (size_t)batch * e->cfg.img_size * e->cfg.img_size * 3;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_img, imgs, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    rc = run_forward(e, batch); if (rc) return rc;
    float* outs[3] = {hs, hm, hl};
    const int C = 3 * (5 + e->cfg.num_classes);
    for (int i = 0; i < 3; i++) {
        if (!outs[i]) continue;
        long long total = (long long)batch * e->g[i] * e->g[i] * C;
        gather_view_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(
            (const float*)e->bufs[e->head_buf[i]].ptr, e->d_user_heads[i], batch, e->g[i], e->g[i], C, e->head_ld, 0);
        e->launches++;
        CUDA_TRY(e, cudaMemcpyAsync(outs[i], e->d_user_heads[i], total * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return Y4_OK;
}

int y4_decode_nms(y4_engine* e, const float* hs, const float* hm, const float* hl, int32_t batch, float iou_thr,
                  float score_thr, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx) {
    int rc = ready(e, batch, false); if (rc) return rc;
    const float* ins[3] = {hs, hm, hl};
    const int C = 3 * (5 + e->cfg.num_classes);
    for (int i = 0; i < 3; i++) {
        if (!ins[i]) return fail(e, Y4_ERR_ARG, "null head tensor");
        CUDA_TRY(e, cudaMemcpyAsync(e->d_user_heads[i], ins[i], sizeof(float) * batch * e->g[i] * e->g[i] * C,
                                    cudaMemcpyHostToDevice, e->stream));
    }
    rc = run_decode_nms(e, batch, iou_thr, score_thr, e->d_user_heads); if (rc) return rc;
    return fetch(e, batch, boxes, scores, classes, valid, cand_idx);
}

int y4_synth_fill(y4_engine* e, uint64_t seed, int64_t first_index, int32_t batch) {
    int rc = ready(e, batch, false); if (rc) return rc;
    const long long per = (long long)e->cfg.img_size * e->cfg.img_size * 3;
    synth_fill_kernel<<<148 * 8, 256, 0, e->stream>>>(e->d_img, seed, (uint64_t)first_index * (uint64_t)per, per * batch);
    e->launches++;
    CUDA_TRY(e, cudaGetLastError());
    return Y4_OK;
}

// Resident path = static launch sequence -> captured once per (batch, part) into a CUDA graph and replayed
// (removes ~110 launch gaps per step).  Y4_GRAPH=0 disables.  what: 1 forward, 2 decode+nms, 3 both.
static int run_resident_part(y4_engine* e, int batch, int what) {
    static const bool use_graph = !(getenv("Y4_GRAPH") && getenv("Y4_GRAPH")[0] == '0');
    auto body = [&]() -> int {
        int rc = Y4_OK;
        if (what & 1) rc = run_forward(e, batch);
        if (!rc && (what & 2)) rc = run_decode_nms(e, batch, e->cfg.iou_threshold, e->cfg.score_threshold, nullptr);
        return rc;
    };
    if (!use_graph) return body();
    const int key = (batch * 4 + what) * 2 + e->img_slot;
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        const int64_t before = e->launches;
        CUDA_TRY(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        int rc = body();
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return fail(e, Y4_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
        cudaGraphExec_t ex = nullptr;
        ce = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail(e, Y4_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
        const int64_t n = e->launches - before;
        e->launches = before;
        it = e->graphs.emplace(key, std::make_pair(ex, n)).first;
    }
    CUDA_TRY(e, cudaGraphLaunch(it->second.first, e->stream));
    e->launches += it->second.second;
    return Y4_OK;
}

int y4_run_forward_resident(y4_engine* e, int32_t batch) {
    int rc = ready(e, batch, true); if (rc) return rc;
    return run_resident_part(e, batch, 1);
}
int y4_run_decode_nms_resident(y4_engine* e, int32_t batch) {
    int rc = ready(e, batch, false); if (rc) return rc;
    return run_resident_part(e, batch, 2);
}
int y4_run_resident(y4_engine* e, int32_t batch) {
    int rc = ready(e, batch, true); if (rc) return rc;
    return run_resident_part(e, batch, 3);
}

int y4_upload_heads(y4_engine* e, const float* hs, const float* hm, const float* hl, int32_t batch) {
    int rc = ready(e, batch, false); if (rc) return rc;
    const float* ins[3] = {hs, hm, hl};
    const int C = 3 * (5 + e->cfg.num_classes);
    for (int i = 0; i < 3; i++) {
        if (!ins[i]) return fail(e, Y4_ERR_ARG, "null head tensor");
        long long total = (long long)batch * e->g[i] * e->g[i] * C;
        CUDA_TRY(e, cudaMemcpyAsync(e->d_user_heads[i], ins[i], total * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        scatter_head_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(
            e->d_user_heads[i], (float*)e->bufs[e->head_buf[i]].ptr, batch, e->g[i], C, e->head_ld, e->d_obj[i], e->obj_rows[i], 5 + e->cfg.num_classes);
        e->launches++;
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return Y4_OK;
}

int y4_fetch_results(y4_engine* e, int32_t batch, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx) {
    int rc = ready(e, batch, false); if (rc) return rc;
    return fetch(e, batch, boxes, scores, classes, valid, cand_idx);
}

// ---- pipelined host path: submit batch i+1 (async H2D on the copy stream) while batch i computes; results come back
// through pinned staging.  Depth 2: at most two submits may be outstanding.
static size_t stage_bytes(const y4_engine* e) {
    const size_t mb = e->cfg.max_boxes, B = e->cfg.max_batch;
    return B * (mb * 4 * 4 + mb * 4 * 3 + 4) + 16;
}
// Source images -> device staging (on `cs`), then the resize + /255 kernel writes the network input of `slot` (on `cs` too,
// so the caller orders the compute stream behind one event).
static int stage_u8(y4_engine* e, int slot, const uint8_t* const* imgs, const int32_t* hs, const int32_t* ws, int batch, int reverse, cudaStream_t cs) {
    if (!imgs || !hs || !ws) return fail(e, Y4_ERR_ARG, "null image table");
    const int S = e->cfg.img_size, B = e->cfg.max_batch;
    size_t total = 0;
    for (int i = 0; i < batch; i++) {
        if (!imgs[i] || hs[i] < 1 || ws[i] < 1 || hs[i] > 32768 || ws[i] > 32768) return fail(e, Y4_ERR_ARG, "bad image in the batch");
        total += ((size_t)hs[i] * ws[i] * 3 + 255) & ~(size_t)255;
    }
    if (!e->div255_ready) {
        float lut[256];
        for (int i = 0; i < 256; i++) lut[i] = (float)((double)i / 255.0);           // img / 255. in float64, then Keras' float32 cast
        CUDA_TRY(e, cudaMemcpyToSymbol(c_div255, lut, sizeof(lut)));
        e->div255_ready = true;
    }
    if (!e->d_pre[slot]) {
        CUDA_TRY(e, cudaMalloc(&e->d_pre[slot], sizeof(PreImage) * B));
        CUDA_TRY(e, cudaHostAlloc((void**)&e->h_pre[slot], sizeof(PreImage) * B, cudaHostAllocDefault));
    }
    if (total > e->u8_cap[slot]) {
        CUDA_TRY(e, cudaStreamSynchronize(cs));
        cudaFree(e->d_u8[slot]); e->d_u8[slot] = nullptr; e->u8_cap[slot] = 0;
        const size_t cap = total + total / 4;
        CUDA_TRY(e, cudaMalloc(&e->d_u8[slot], cap));
        e->u8_cap[slot] = cap;
    }
    size_t off = 0;
    for (int i = 0; i < batch; i++) {
        const size_t n = (size_t)hs[i] * ws[i] * 3;
        e->h_pre[slot][i] = PreImage{(long long)off, hs[i], ws[i]};
        CUDA_TRY(e, cudaMemcpyAsync(e->d_u8[slot] + off, imgs[i], n, cudaMemcpyHostToDevice, cs));
        off += (n + 255) & ~(size_t)255;
    }
    CUDA_TRY(e, cudaMemcpyAsync(e->d_pre[slot], e->h_pre[slot], sizeof(PreImage) * batch, cudaMemcpyHostToDevice, cs));
    dim3 grid((unsigned)((S + 255) / 256), (unsigned)S, (unsigned)batch);
    preprocess_u8_kernel<<<grid, 256, 0, cs>>>(e->d_u8[slot], e->d_pre[slot], e->d_img_slot[slot], S, batch, reverse);
    e->launches++;
    CUDA_TRY(e, cudaGetLastError());
    return Y4_OK;
}

int y4_preprocess_u8(y4_engine* e, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths, int32_t batch,
                     int32_t reverse_channels, float* out) {
    int rc = ready(e, batch, false); if (rc) return rc;
    rc = stage_u8(e, 0, imgs, heights, widths, batch, reverse_channels, e->stream); if (rc) return rc;
    if (out) CUDA_TRY(e, cudaMemcpyAsync(out, e->d_img_slot[0], sizeof(float) * 3 * e->cfg.img_size * e->cfg.img_size * batch, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return Y4_OK;
}

int y4_predict_u8(y4_engine* e, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths, int32_t batch,
                  int32_t reverse_channels, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx) {
    int rc = ready(e, batch, true); if (rc) return rc;
    rc = stage_u8(e, 0, imgs, heights, widths, batch, reverse_channels, e->stream); if (rc) return rc;
    rc = run_resident_part(e, batch, 3); if (rc) return rc;
    return fetch(e, batch, boxes, scores, classes, valid, cand_idx);
}

static int submit_common(y4_engine* e, int32_t batch, const float* imgs, const uint8_t* const* u8, const int32_t* hs, const int32_t* ws, int reverse);

int y4_submit_u8(y4_engine* e, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths, int32_t batch, int32_t reverse_channels) {
    return submit_common(e, batch, nullptr, imgs, heights, widths, reverse_channels);
}

int y4_submit(y4_engine* e, const float* imgs, int32_t batch) {
    if (!imgs) return e ? fail(e, Y4_ERR_ARG, "null imgs") : Y4_ERR_ARG;
    return submit_common(e, batch, imgs, nullptr, nullptr, nullptr, 0);
}

static int submit_common(y4_engine* e, int32_t batch, const float* imgs, const uint8_t* const* u8, const int32_t* hs, const int32_t* ws, int reverse) {
    int rc = ready(e, batch, true, true); if (rc) return rc;
    if (e->n_submitted - e->n_collected >= 2) return fail(e, Y4_ERR_STATE, "two batches already in flight: call y4_collect first");
    const int S = e->cfg.img_size, B = e->cfg.max_batch, mb = e->cfg.max_boxes;
    if (!e->copy_stream) {
        CUDA_TRY(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(e, cudaMalloc(&e->d_img_slot[1], sizeof(float) * 3 * S * S * B));
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
            CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
            CUDA_TRY(e, cudaHostAlloc((void**)&e->stage[i], stage_bytes(e), cudaHostAllocDefault));
        }
    }
    const int slot = (int)(e->n_submitted & 1);
    // the slot's previous occupant (submit n-2) was collected, hence its compute (which read d_img_slot[slot]) is done
    const size_t n = (size_t)batch * S * S * 3;
    if (imgs) CUDA_TRY(e, cudaMemcpyAsync(e->d_img_slot[slot], imgs, n * sizeof(float), cudaMemcpyHostToDevice, e->copy_stream));
    else { rc = stage_u8(e, slot, u8, hs, ws, batch, reverse, e->copy_stream); if (rc) return rc; }
    CUDA_TRY(e, cudaEventRecord(e->ev_h2d[slot], e->copy_stream));
    CUDA_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_h2d[slot], 0));
    e->d_img = e->d_img_slot[slot]; e->img_slot = slot;
    rc = run_resident_part(e, batch, 3);
    e->d_img = e->d_img_slot[0]; e->img_slot = 0;
    if (rc) return rc;
    char* st = e->stage[slot];
    size_t off = 0;
    CUDA_TRY(e, cudaMemcpyAsync(st + off, e->d_out_boxes, sizeof(float) * 4 * mb * batch, cudaMemcpyDeviceToHost, e->stream)); off += sizeof(float) * 4 * mb * B;
    CUDA_TRY(e, cudaMemcpyAsync(st + off, e->d_out_scores, sizeof(float) * mb * batch, cudaMemcpyDeviceToHost, e->stream)); off += sizeof(float) * mb * B;
    CUDA_TRY(e, cudaMemcpyAsync(st + off, e->d_out_classes, sizeof(float) * mb * batch, cudaMemcpyDeviceToHost, e->stream)); off += sizeof(float) * mb * B;
    CUDA_TRY(e, cudaMemcpyAsync(st + off, e->d_out_idx, sizeof(int) * mb * batch, cudaMemcpyDeviceToHost, e->stream)); off += sizeof(int) * mb * B;
    CUDA_TRY(e, cudaMemcpyAsync(st + off, e->d_out_valid, sizeof(int) * batch, cudaMemcpyDeviceToHost, e->stream)); off += sizeof(int) * B;
    CUDA_TRY(e, cudaEventRecord(e->ev_done[slot], e->stream));
    e->sub_batch[slot] = batch;
    e->n_submitted++;
    return Y4_OK;
}

int y4_collect(y4_engine* e, int32_t batch, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx) {
    if (!e) return Y4_ERR_ARG;
    if (e->n_collected >= e->n_submitted) return fail(e, Y4_ERR_STATE, "nothing submitted");
    const int slot = (int)(e->n_collected & 1);
    if (batch != e->sub_batch[slot]) return fail(e, Y4_ERR_ARG, "batch differs from the submitted one");
    cudaSetDevice(e->cfg.device);
    CUDA_TRY(e, cudaEventSynchronize(e->ev_done[slot]));
    e->n_collected++;
    const size_t mb = e->cfg.max_boxes, B = e->cfg.max_batch;
    const char* st = e->stage[slot];
    size_t off = 0;
    if (boxes) memcpy(boxes, st + off, sizeof(float) * 4 * mb * batch);
    off += sizeof(float) * 4 * mb * B;
    if (scores) memcpy(scores, st + off, sizeof(float) * mb * batch);
    off += sizeof(float) * mb * B;
    if (classes) memcpy(classes, st + off, sizeof(float) * mb * batch);
    off += sizeof(float) * mb * B;
    if (cand_idx) memcpy(cand_idx, st + off, sizeof(int) * mb * batch);
    off += sizeof(int) * mb * B;
    if (valid) memcpy(valid, st + off, sizeof(int) * batch);
    return Y4_OK;
}

int y4_sync(y4_engine* e) {
    if (!e) return Y4_ERR_ARG;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return Y4_OK;
}
int y4_timer_begin(y4_engine* e) {
    if (!e) return Y4_ERR_ARG;
    CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
    return Y4_OK;
}
int y4_timer_end(y4_engine* e, float* ms) {
    if (!e || !ms) return Y4_ERR_ARG;
    CUDA_TRY(e, cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(e, cudaEventSynchronize(e->ev1));
    CUDA_TRY(e, cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return Y4_OK;
}
int y4_flush_l2(y4_engine* e) {
    if (!e) return Y4_ERR_ARG;
    l2_flush_kernel<<<148 * 8, 256, 0, e->stream>>>(e->d_flush, (long long)e->flush_elems, 1.0f);
    CUDA_TRY(e, cudaGetLastError());
    return Y4_OK;
}
int64_t y4_launch_count(const y4_engine* e) { return e ? e->launches : 0; }

int y4_profile_layers(y4_engine* e, int32_t batch, float* ms, int32_t n) {
    int rc = ready(e, batch, true); if (rc) return rc;
    if (!ms || n < (int)e->steps.size()) return fail(e, Y4_ERR_ARG, "ms buffer too small");
    std::vector<cudaEvent_t> ev(e->steps.size() + 1);
    for (auto& x : ev) cudaEventCreate(&x);
    cudaEventRecord(ev[0], e->stream);
    for (size_t i = 0; i < e->steps.size(); i++) {
        auto& s = e->steps[i];
        rc = run_step(e, s, batch); if (rc) return rc;
        cudaEventRecord(ev[i + 1], e->stream);
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < e->steps.size(); i++) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    for (auto& x : ev) cudaEventDestroy(x);
    return (int)e->steps.size();
}

void* y4_host_alloc(size_t nbytes) { void* p = nullptr; return cudaHostAlloc(&p, nbytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
void y4_host_free(void* p) { if (p) cudaFreeHost(p); }

int y4_num_layers(const y4_engine* e) {                 // the 110 convs of the reference (fused sibling entries come after them)
    if (!e) return 0;
    int n = 0;
    for (auto& c : e->convs) n += c.fused_a < 0;
    return n;
}
int y4_num_steps(const y4_engine* e) { return e ? (int)e->steps.size() : 0; }
int64_t y4_num_boxes(const y4_engine* e) { return e ? e->N : 0; }

int y4_describe_layer(const y4_engine* e, int32_t idx, y4_layer_info* info) {
    if (!e || !info || idx < 0 || idx >= (int)e->convs.size()) return Y4_ERR_ARG;
    const ConvOp& c = e->convs[idx];
    info->idx = c.idx; info->cin = c.cin; info->cout = c.cout; info->ksize = c.k; info->stride = c.stride;
    info->batch_norm = c.bn; info->activation = c.act; info->out_hw = c.N_OH; info->kernel_kind = c.kind;
    info->tile_n = c.kind ? c.tc.tile_n : 0;
    info->flops = 2ll * c.N_OH * c.N_OH * c.cout * c.K;
    snprintf(info->out_name, sizeof(info->out_name), "%s", c.out_name.c_str());
    const bool tc = c.kind == 1 || c.kind == 2;
    info->tc_mode = tc ? (c.tc.cta2 ? (c.tc.p.mode == 3 ? 5 : 4) : c.tc.p.mode) : 0; info->tc_epilogue = tc && c.tc.p.epi ? c.tc.p.epi_gw : 0; info->tc_stages = tc ? c.tc.stages : 0;
    info->tc_group = tc ? c.tc.p.group : 0; info->tc_ctas_per_sm = tc ? c.tc.ctas_per_sm : 0; info->tc_bk = tc ? c.tc.bk : 0;
    info->tc_epi_warps = tc ? (c.tc.lean ? 44 : c.tc.nepi) : 0; info->tc_resident_w = tc ? c.tc.p.bres : 0;
    return Y4_OK;
}

int y4_describe_step(const y4_engine* e, int32_t step, y4_layer_info* info) {
    if (!e || !info || step < 0 || step >= (int)e->steps.size()) return Y4_ERR_ARG;
    const Step& st = e->steps[step];
    if (st.type == 0) {
        int rc = y4_describe_layer(e, st.conv, info);
        if (rc) return rc;
        const ConvOp& c = e->convs[st.conv];
        if (c.fused_a >= 0) info->flops = 2ll * c.N_OH * c.N_OH * c.cout * c.K;
        if (c.chain >= 0) {
            const ConvOp& q = e->convs[c.chain];
            info->flops += 2ll * q.N_OH * q.N_OH * q.cout * q.K;
            snprintf(info->out_name, sizeof(info->out_name), "%s>%s", c.out_name.c_str(), q.out_name.c_str());
        }
        return Y4_OK;
    }
    memset(info, 0, sizeof(*info));
    if (st.type == 3) {                                     // conv 0 + conv 1 in one kernel
        const ConvOp &c0 = e->convs[0], &c1 = e->convs[1];
        info->idx = -2; info->kernel_kind = 6; info->cin = c0.cin; info->cout = c1.cout; info->ksize = 3; info->stride = 2;
        info->batch_norm = 1; info->activation = c1.act; info->out_hw = c1.N_OH; info->tile_n = 64;
        info->flops = 2ll * c0.N_OH * c0.N_OH * c0.cout * c0.K + 2ll * c1.N_OH * c1.N_OH * c1.cout * c1.K;
        snprintf(info->out_name, sizeof(info->out_name), "c0+c1");
        return Y4_OK;
    }
    info->idx = -1; info->kernel_kind = 5;                  // SPP max-pools
    info->out_hw = e->bufs[e->spp_view.buf].H; info->cin = info->cout = e->spp_C;
    snprintf(info->out_name, sizeof(info->out_name), "spp");
    return Y4_OK;
}

int y4_debug_run_conv(y4_engine* e, int32_t idx, int32_t batch, int32_t use_tc) {
    int rc = ready(e, batch, true); if (rc) return rc;
    if (idx < 0 || idx >= (int)e->convs.size()) return fail(e, Y4_ERR_ARG, "bad conv idx");
    const ConvOp& c = e->convs[idx];
    if (use_tc) {
        if (c.kind == 0) return fail(e, Y4_ERR_ARG, "conv has no tcgen05 plan");
        rc = run_conv(e, c, batch); if (rc) return rc;
    } else {
        if (e->cfg.precision == Y4_PREC_FP16X3) return fail(e, Y4_ERR_ARG, "no CUDA-core kernel for split-precision buffers");
        if (c.fused_a >= 0) return fail(e, Y4_ERR_ARG, "fused sibling convs have no CUDA-core kernel: run the two convs");
        launch_simt(e, c, batch);
    }
    CUDA_TRY(e, cudaGetLastError());
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return Y4_OK;
}

// Phase timestamps of the tcgen05 conv kernel (clock64 per CTA): out[cta*16 + {0 entry,1 setup done,2 first TMA issued,
// 3 last TMA issued,4 first stage landed,5 last MMA issued,6 accumulator ready,7 epilogue done,8 exit,14 globaltimer,15 smid}]
int y4_debug_trace_conv(y4_engine* e, int32_t idx, int32_t batch, int64_t* out, int32_t max_ctas) {
    int rc = ready(e, batch, true); if (rc) return rc;
    if (idx < 0 || idx >= (int)e->convs.size() || !out || max_ctas < 1) return fail(e, Y4_ERR_ARG, "bad args");
    const ConvOp& c = e->convs[idx];
    if (c.kind != 1 && c.kind != 2) return fail(e, Y4_ERR_ARG, "conv has no tcgen05 plan");
    if (max_ctas > 4096) max_ctas = 4096;
    long long* d = nullptr;
    CUDA_TRY(e, cudaMalloc(&d, sizeof(long long) * 16 * 4096));
    CUDA_TRY(e, cudaMemsetAsync(d, 0, sizeof(long long) * 16 * 4096, e->stream));
    TcConvPlan pl = c.tc;
    tc_launch(pl, batch, e->stream);                 // warm
    pl.p.dbg = d;
    int lr = tc_launch(pl, batch, e->stream);
    cudaError_t ce = cudaMemcpyAsync(out, d, sizeof(long long) * 16 * max_ctas, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(d);
    if (lr != 0 || ce != cudaSuccess) return fail(e, Y4_ERR_CUDA, "trace launch failed");
    return Y4_OK;
}

int64_t y4_debug_get_tensor(y4_engine* e, const char* name, int32_t batch, float* out, int64_t capacity) {
    if (!e || !name) return Y4_ERR_ARG;
    int rc = check_batch(e, batch); if (rc) return rc;
    auto it = e->views.find(name);
    if (it == e->views.end() || it->second.buf < 0) return fail(e, Y4_ERR_ARG, std::string("unknown tensor ") + name);
    const View& v = it->second;
    const Buf& b = e->bufs[v.buf];
    long long total = (long long)batch * v.H * v.W * v.C;
    if (!out) return total;
    if (capacity < total) return fail(e, Y4_ERR_ARG, "capacity too small");
    cudaSetDevice(e->cfg.device);
    float* tmp = nullptr;
    CUDA_TRY(e, cudaMalloc(&tmp, total * sizeof(float)));
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (b.elt == 4) gather_view_kernel<float><<<blocks, 256, 0, e->stream>>>((const float*)b.ptr, tmp, batch, v.H, v.W, v.C, b.C, v.choff);
    else if (b.ptr_lo) gather_view_split_kernel<<<blocks, 256, 0, e->stream>>>((const __half*)b.ptr, (const __half*)b.ptr_lo, tmp, batch, v.H, v.W, v.C, b.C, v.choff, 1.0f / 256.0f);
    else gather_view_kernel<__half><<<blocks, 256, 0, e->stream>>>((const __half*)b.ptr, tmp, batch, v.H, v.W, v.C, b.C, v.choff);
    cudaError_t err = cudaMemcpyAsync(out, tmp, total * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(tmp);
    if (err != cudaSuccess) return fail(e, Y4_ERR_CUDA, cudaGetErrorString(err));
    return total;
}

// ---- multi-GPU ------------------------------------------------------------------------------------
int y4_comm_unique_id(void* uid128) {
    if (!uid128) return Y4_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (!nccl().ok) return fail(nullptr, Y4_ERR_COMM, "libnccl.so.2 could not be loaded");
    if (nccl().GetUniqueId(&id) != ncclSuccess) return Y4_ERR_COMM;
    memcpy(uid128, &id, 128);
    return Y4_OK;
}

int y4_comm_init(y4_engine* e, int32_t rank, int32_t nranks, const void* uid128) {
    if (!e || !uid128 || nranks < 1 || rank < 0 || rank >= nranks) return e ? fail(e, Y4_ERR_ARG, "bad comm args") : Y4_ERR_ARG;
    cudaSetDevice(e->cfg.device);
    ncclUniqueId id; memcpy(&id, uid128, 128);
    if (!nccl().ok) return fail(e, Y4_ERR_COMM, "libnccl.so.2 could not be loaded");
    ncclResult_t r = nccl().CommInitRank(&e->comm, nranks, id, rank);
    if (r != ncclSuccess) return fail(e, Y4_ERR_COMM, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    e->rank = rank; e->nranks = nranks;
    const int mb = e->cfg.max_boxes;
    e->gather_bytes = (size_t)nranks * e->cfg.max_batch * (sizeof(float) * 6 * mb + sizeof(int) * (mb + 1));
    CUDA_TRY(e, cudaMalloc(&e->d_gather, e->gather_bytes));
    return Y4_OK;
}

int y4_allgather_results(y4_engine* e, int32_t batch, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx) {
    int rc = ready(e, batch, false); if (rc) return rc;
    if (!e->comm) return fail(e, Y4_ERR_STATE, "y4_comm_init not called");
    const int mb = e->cfg.max_boxes, R = e->nranks;
    char* g = static_cast<char*>(e->d_gather);
    float* gb = reinterpret_cast<float*>(g);                       g += sizeof(float) * 4 * mb * batch * R;
    float* gs = reinterpret_cast<float*>(g);                       g += sizeof(float) * mb * batch * R;
    float* gc = reinterpret_cast<float*>(g);                       g += sizeof(float) * mb * batch * R;
    int* gv = reinterpret_cast<int*>(g);                           g += sizeof(int) * batch * R;
    int* gi = reinterpret_cast<int*>(g);
    ncclResult_t r = nccl().GroupStart();
    if (r == ncclSuccess) r = nccl().AllGather(e->d_out_boxes, gb, (size_t)4 * mb * batch, ncclFloat, e->comm, e->stream);
    if (r == ncclSuccess) r = nccl().AllGather(e->d_out_scores, gs, (size_t)mb * batch, ncclFloat, e->comm, e->stream);
    if (r == ncclSuccess) r = nccl().AllGather(e->d_out_classes, gc, (size_t)mb * batch, ncclFloat, e->comm, e->stream);
    if (r == ncclSuccess) r = nccl().AllGather(e->d_out_valid, gv, (size_t)batch, ncclInt32, e->comm, e->stream);
    if (r == ncclSuccess) r = nccl().AllGather(e->d_out_idx, gi, (size_t)mb * batch, ncclInt32, e->comm, e->stream);
    if (r == ncclSuccess) r = nccl().GroupEnd();
    if (r != ncclSuccess) return fail(e, Y4_ERR_COMM, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
    const size_t nb = (size_t)batch * R;
    if (boxes) CUDA_TRY(e, cudaMemcpyAsync(boxes, gb, sizeof(float) * 4 * mb * nb, cudaMemcpyDeviceToHost, e->stream));
    if (scores) CUDA_TRY(e, cudaMemcpyAsync(scores, gs, sizeof(float) * mb * nb, cudaMemcpyDeviceToHost, e->stream));
    if (classes) CUDA_TRY(e, cudaMemcpyAsync(classes, gc, sizeof(float) * mb * nb, cudaMemcpyDeviceToHost, e->stream));
    if (valid) CUDA_TRY(e, cudaMemcpyAsync(valid, gv, sizeof(int) * nb, cudaMemcpyDeviceToHost, e->stream));
    if (cand_idx) CUDA_TRY(e, cudaMemcpyAsync(cand_idx, gi, sizeof(int) * mb * nb, cudaMemcpyDeviceToHost, e->stream));
    if (boxes || scores || classes || valid || cand_idx) CUDA_TRY(e, cudaStreamSynchronize(e->stream));   // all-NULL: device-side gather only, async
    return Y4_OK;
}

}  // extern "C"
