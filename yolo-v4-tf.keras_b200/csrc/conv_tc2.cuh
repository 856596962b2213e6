// cta_group::2 variant of the flat tcgen05 conv kernel (conv_tc.cuh): a cluster of two CTAs on an SM pair computes a
// 256-row x BN tile with ONE tcgen05.mma.cta_group::2 per K step.
//
// Why: with SS operands a 128 x N x 16 MMA reads its A (4 KB) and B (N * 32 B) slices from shared memory for every
// instruction.  At N = 128 that is 8 KB per 64 tensor cycles = the whole 128 B/clk shared-memory port, before the TMA has
// written a byte; at N = 256 it is 96 B/clk, and the TMA writes as much again (measured: 45 % / 60 % tensor-pipe active,
// MMA thread waiting on data 17-20 %).  In a CTA pair each CTA stages its own 128 rows of A and only HALF of the B tile
// (BN/2 weight rows); the hardware feeds both halves to both tensor cores, so B traffic per SM (L2 -> smem and
// smem -> tensor core) halves.
//
// Roles per CTA as in conv_tc.cuh (warp 0 TMA producer, warp 1 MMA / TMEM alloc, warps 2.. epilogue), except:
//   * every TMA load names the LEADER's (cluster rank 0) full barrier: one barrier completes when both CTAs' bytes landed;
//   * only the leader issues MMAs; tcgen05.commit multicasts the arrive to both CTAs' empty / tfull barriers;
//   * the peer's epilogue warps arrive remotely on the leader's tempty barrier (mbarrier.arrive.shared::cluster);
//   * TMEM is allocated with cta_group::2 by warp 1 of both CTAs; cluster barriers fence start-up and tear-down.
// Flat mode (1x1 and 3x3 stride 1; slab epilogue) and strided-box mode (3x3 stride 2: each CTA of the pair owns one TH x TW
// output box, A comes from the four parity planes as in conv_tc.cuh, per-thread stores); 64-channel K blocks, fp16 in / out.
#pragma once
#include "conv_tc.cuh"

namespace y4 {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;              // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// elect-predicated issue for warp-uniform loops (see elect_one in conv_tc.cuh)
__device__ __forceinline__ void umma_f16_cg2_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_elect(uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
        "}" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {       // from either CTA of the pair
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

template <int BN, int NEPI>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1) conv_tc2_kernel(const __grid_constant__ TcParams p) {
    constexpr int BK = 64, SWZ = 128;
    constexpr int A_BYTES = 128 * BK * 2;
    constexpr int B_BYTES = (BN / 2) * BK * 2;              // this CTA's half of the weight tile
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;                  // two accumulator stages
    constexpr uint32_t IDESC = make_idesc(256, BN);

    extern __shared__ unsigned char tc_smem[];
    const uint32_t raw = smem_u32(tc_smem);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const int S = p.stages, G = p.group;
    // mode 3 (3x3 stride 1, A-patch reuse): [patch_slots patches][S stages of this CTA's half B tile]; otherwise S*G stages of (A | B)
    const bool patch = p.mode == 3;
    const uint32_t ring_bytes = patch ? (uint32_t)(p.patch_slots * p.patch_bytes) + (uint32_t)S * (uint32_t)B_BYTES
                                      : (uint32_t)(S * G) * (uint32_t)STAGE_BYTES;
    const uint32_t bring = base + (uint32_t)(p.patch_slots * p.patch_bytes);   // mode 3: first B stage
    const uint32_t epi_bytes = p.epi ? epi_slab_bytes(NEPI, p.epi_gw) : 0u;
    // chain fusion: A2 = this CTA's 128 output rows x 64*q_kb channels as q_kb K-major SWIZZLE_128B blocks (the second MMA's A
    // operand, written by the epilogue warps); W2 = this CTA's half of Q's weights, resident
    const uint32_t a2_bytes = p.q_on ? (uint32_t)p.q_kb * 16384u : 0u;
    const uint32_t w2_blk = (uint32_t)(p.q_n / 2) * 128u;
    const uint32_t w2_bytes = p.q_on ? (uint32_t)p.q_kb * w2_blk : 0u;
    const uint32_t a2 = base + ring_bytes, w2 = a2 + a2_bytes;
    const uint32_t slabs = w2 + ((w2_bytes + 1023u) & ~1023u);
    const uint32_t bars = slabs + epi_bytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8u * S, bar_tfull = bars + 16u * S, bar_tempty = bars + 16u * S + 16u;
    const uint32_t tmem_slot = bars + 16u * S + 32u;
    const uint32_t bar_pfull = bars + 16u * S + 48u, bar_pempty = bars + 16u * S + 112u;   // 8 patch slots each (as conv_tc.cuh)
    const uint32_t bar_a2full = bars + 16u * S + 176u, bar_t2full = bars + 16u * S + 184u, bar_w2 = bars + 16u * S + 192u;
    float* sbias = reinterpret_cast<float*>(tc_smem + (slabs - raw) + epi_bytes + 16u * S + 224u);
    float* sbias2 = sbias + p.bias_n;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = (int)(blockIdx.x >> 1), nclusters = (int)(gridDim.x >> 1);
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmW);
        tma_prefetch_desc(&p.tmA[0]);
        if (p.epi) { tma_prefetch_desc(&p.tmOut); if (p.split_col) tma_prefetch_desc(&p.tmOut2); }
        if (p.mode == 2) { tma_prefetch_desc(&p.tmA[1]); tma_prefetch_desc(&p.tmA[2]); tma_prefetch_desc(&p.tmA[3]); }
        for (int s = 0; s < S; s++) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(bar_tfull + 8u * a, 1); mbar_init(bar_tempty + 8u * a, 2 * NEPI); }
        for (int a = 0; a < 8; a++) { mbar_init(bar_pfull + 8u * a, 1); mbar_init(bar_pempty + 8u * a, 1); }
        mbar_init(bar_a2full, 2 * NEPI); mbar_init(bar_t2full, 1); mbar_init(bar_w2, 1);
        if (p.q_on) { tma_prefetch_desc(&p.tmW2); tma_prefetch_desc(&p.tmOutQ); }
        if (patch) tma_prefetch_desc(&p.tmA[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_cg2(tmem_slot, TMEM_COLS);
    const float bmul = p.act >= 2 ? 1.4426950408889634f : 1.0f;          // mish layers keep b * log2(e) (act_fast)
    if (warp >= 2) for (int i = threadIdx.x - 64; i < p.bias_n; i += 32 * NEPI) sbias[i] = p.bias[i] * bmul;
    if (warp >= 2 && p.q_on) for (int i = threadIdx.x - 64; i < p.q_n; i += 32 * NEPI) sbias2[i] = p.q_bias[i] * (p.q_act == 2 ? 1.4426950408889634f : 1.0f);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                     // the peer's barriers exist before anything is signalled across the pair
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();

    // cluster tile ct -> (256-row block, N tile); this CTA: rows +128*rank, weight rows +BN/2*rank
    auto tile_m0 = [&](int ct) { return (long long)p.m_start + (long long)(ct / p.n_tiles) * 256 + 128 * (long long)rank; };
    auto tile_n0 = [&](int ct) { return (ct % p.n_tiles) * BN; };
    // box mode: M tile index of this CTA (the odd CTA of the last pair may have none: it repeats the last tile and stores nothing)
    const int m_tiles_box = p.mode == 2 ? (int)(p.M_total) : 0;                 // tc2_launch: batch * tiles_per_img
    auto box_tile = [&](int ct, int& img, int& oh0, int& ow0) {
        int mt = 2 * (ct / p.n_tiles) + (int)rank;
        const bool real = mt < m_tiles_box;
        if (!real) mt = m_tiles_box - 1;
        img = mt / p.tiles_per_img;
        const int r = mt - img * p.tiles_per_img;
        const int th = r / p.tiles_w;
        oh0 = th * p.TH; ow0 = (r - th * p.tiles_w) * p.TW;
        return real;
    };

    if (warp == 0) {
        if (p.q_on) {
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(bar_w2, 2u * w2_bytes);
                for (int kb2 = 0; kb2 < p.q_kb; kb2++)
                    tma_load_2d_cg2(w2 + (uint32_t)kb2 * w2_blk, &p.tmW2, bar_w2 & kPeerBitMask, kb2 * BK, (int)rank * (p.q_n / 2));
            }
            __syncwarp();
        }
        if (patch) {
            // A-patch reuse: per (tile, 64-channel block) ONE patch of 130 + 2*Wp rows (this CTA's 128 output rows shifted by
            // -(Wp+1) .. +(Wp+1)) feeds all nine taps through row-shifted UMMA descriptors; only the weights ride the ring: the A
            // traffic from L2 drops to (130 + 2*Wp) / (9 * 128) of the per-tap loads.  Same K order (channel block outer, tap inner).
            // All 32 lanes run the loop (warp-uniform, see elect_one); patches are issued up to PS - 2 ahead of the weight loads.
            const int ncb = p.kb_per_tap;
            const int my_tiles = (p.num_tiles - cluster_id + nclusters - 1) / nclusters;
            const int npatch = my_tiles > 0 ? my_tiles * ncb : 0;
            const uint32_t PS = (uint32_t)p.patch_slots;
            const uint32_t ptx = (uint32_t)p.patch_boxes * (uint32_t)p.patch_box_rows * 128u;
            uint32_t bs = 0, bph = 0, ps = 0, pph = 0;       // weight-ring stage / phase, patch slot / phase
            int issued = 0, itl = 0, icb = 0;                // next patch to issue: its tile (local index) and channel block
            const int ahead = (int)PS > 2 ? (int)PS - 2 : 1;
            int tl = 0, cb = 0;
            for (int j = 0; j < npatch; j++) {
                while (issued < npatch && issued <= j + ahead) {
                    const long long m0 = tile_m0(cluster_id + itl * nclusters);
                    mbar_wait(bar_pempty + 8u * ps, pph ^ 1u);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(bar_pfull + 8u * ps, 2u * ptx);
                        const uint32_t fb = (bar_pfull + 8u * ps) & kPeerBitMask;
                        const uint32_t dst = base + ps * (uint32_t)p.patch_bytes;
                        const int row0 = (int)m0 - 1 - p.Wp;
                        for (int b = 0; b < p.patch_boxes; b++)
                            tma_load_2d_cg2(dst + (uint32_t)(b * p.patch_box_rows) * 128u, &p.tmA[1], fb, icb * BK, row0 + b * p.patch_box_rows);
                    }
                    __syncwarp();
                    if (++ps == PS) { ps = 0; pph ^= 1u; }
                    if (++icb == ncb) { icb = 0; itl++; }
                    issued++;
                }
                const int n0 = tile_n0(cluster_id + tl * nclusters);
                for (int tap = 0; tap < 9; tap++) {
                    mbar_wait(bar_empty + 8u * bs, bph ^ 1u);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(bar_full + 8u * bs, 2u * (uint32_t)B_BYTES);
                        tma_load_2d_cg2(bring + bs * (uint32_t)B_BYTES, &p.tmW, (bar_full + 8u * bs) & kPeerBitMask, (tap * ncb + cb) * BK,
                                        n0 + (int)rank * (BN / 2));
                    }
                    __syncwarp();
                    if (++bs == (uint32_t)S) { bs = 0; bph ^= 1u; }
                }
                if (++cb == ncb) { cb = 0; tl++; }
            }
        } else if (!patch) {
            // all 32 lanes run the loop (warp-uniform control flow, see elect_one); one elected lane issues
            uint32_t s = 0, ph = 0;                                                   // ring stage and its phase (no per-stage division)
            for (int ct = cluster_id; ct < p.num_tiles; ct += nclusters) {
                const long long m0 = tile_m0(ct);
                const int n0 = tile_n0(ct);
                int bimg = 0, boh0 = 0, bow0 = 0;
                if (p.mode == 2) box_tile(ct, bimg, boh0, bow0);
                const uint32_t a_bytes = p.mode == 2 ? (uint32_t)(p.TH * p.TW * BK * 2) : (uint32_t)A_BYTES;
                for (int kb0 = 0; kb0 < p.num_kb; kb0 += G) {
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    const uint32_t fb = (bar_full + 8u * s) & kPeerBitMask;          // the leader's barrier collects both CTAs' bytes
                    const int gcount = p.num_kb - kb0 < G ? p.num_kb - kb0 : G;
                    if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(bar_full + 8u * s, 2u * (uint32_t)gcount * (a_bytes + (uint32_t)B_BYTES));
                    for (int kk = 0; kk < gcount; kk++) {
                        const int kb = kb0 + kk;
                        const uint32_t sa = base + (s * (uint32_t)G + (uint32_t)kk) * (uint32_t)STAGE_BYTES;
                        int tap = 0, cb = kb;
                        if (p.mode == 2) {                                            // box: tap outer (as conv_tc.cuh mode 2)
                            tap = kb / p.kb_per_tap; cb = kb - tap * p.kb_per_tap;
                            if (p.pairx) {                                            // pixel-pair view (conv_tc.cuh TcConvDesc::pairx)
                                const int kh = tap >> 1, j = tap & 1;
                                tma_load_4d_cg2(sa, &p.tmA[(kh & 1) * 2], fb, cb * BK, bow0 + j, boh0 + (kh >> 1), bimg);
                            } else {
                                const int kh = tap / 3, kw = tap - kh * 3;
                                tma_load_4d_cg2(sa, &p.tmA[(kh & 1) * 2 + (kw & 1)], fb, cb * BK, bow0 + (kw >> 1), boh0 + (kh >> 1), bimg);
                            }
                        } else {
                            if (p.ksize == 3) { cb = kb / 9; tap = kb - cb * 9; }    // channel block outer, tap inner (as conv_tc.cuh)
                            int shift = 0;
                            if (p.ksize == 3) { const int kh = tap / 3, kw = tap - kh * 3; shift = (kh - 1) * p.Wp + (kw - 1); }
                            tma_load_2d_cg2(sa, &p.tmA[0], fb, cb * BK, (int)(m0 + shift));
                        }
                        tma_load_2d_cg2(sa + A_BYTES, &p.tmW, fb, (tap * p.kb_per_tap + cb) * BK, n0 + (int)rank * (BN / 2));
                    }
                    }
                    __syncwarp();
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                                     // all 32 lanes run the loop; the MMAs / commits are elect-predicated
            uint32_t ti = 0, ms = 0, mph = 0, mps = 0, mpph = 0;
            // chain fusion: the second MMA (A2 x W2 -> the TMEM stage the epilogue has just drained) of the PREVIOUS tile is issued
            // from inside this tile's K loop as soon as the epilogue warps of both CTAs have written A2 (non-blocking test per stage),
            // at the latest at the end of the K loop: the tensor pipe never waits for the epilogue, and the stage is handed back
            // (second epilogue) before the tile after this one needs it.
            bool pend = false, w2_ready = false;
            uint32_t pend_as = 0, q_ph = 0;
            const uint32_t IDESC2 = make_idesc(256, p.q_on ? p.q_n : BN);
            auto issue_m2 = [&]() {
                if (!w2_ready) { mbar_wait(bar_w2, 0u); w2_ready = true; }
                tc_fence_after();
                const uint32_t tacc2 = tmem_base + pend_as * (uint32_t)BN;
                for (int kb2 = 0; kb2 < p.q_kb; kb2++) {
                    const uint64_t da = make_smem_desc<SWZ>(a2 + (uint32_t)kb2 * 16384u);
                    const uint64_t db = make_smem_desc<SWZ>(w2 + (uint32_t)kb2 * w2_blk);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        umma_f16_cg2_elect(tacc2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC2, (kb2 | k) ? 1u : 0u);
                }
                umma_commit_cg2_elect(bar_t2full);
                pend = false; q_ph ^= 1u;
            };
            for (int ct = cluster_id; ct < p.num_tiles; ct += nclusters, ti++) {
                const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
                mbar_wait(bar_tempty + 8u * as, aph ^ 1u);                           // both CTAs' epilogues have drained this stage
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * (uint32_t)BN;
                if (patch) {
                    const uint32_t PS = (uint32_t)p.patch_slots;
                    for (int cb = 0; cb < p.kb_per_tap; cb++) {
                        const uint32_t ps = mps, pph = mpph;
                        if (++mps == PS) { mps = 0; mpph ^= 1u; }
                        mbar_wait(bar_pfull + 8u * ps, pph);
                        tc_fence_after();
                        const uint32_t pa = base + ps * (uint32_t)p.patch_bytes;
                        for (int tap = 0; tap < 9; tap++) {
                            const uint32_t s = ms, ph = mph;
                            if (++ms == (uint32_t)S) { ms = 0; mph ^= 1u; }
                            if (pend && mbar_test(bar_a2full, q_ph)) issue_m2();
                            mbar_wait(bar_full + 8u * s, ph);
                            tc_fence_after();
                            // patch row 0 is output row m0 shifted by -(Wp+1): tap (kh, kw) starts at row kh*Wp + kw
                            const int roff = (tap / 3) * p.Wp + (tap % 3);
                            const uint64_t da = make_smem_desc<SWZ>(pa + (uint32_t)roff * 128u);
                            const uint64_t db = make_smem_desc<SWZ>(bring + s * (uint32_t)B_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / 16; k++)
                                umma_f16_cg2_elect(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (cb | tap | k) ? 1u : 0u);
                            umma_commit_cg2_elect(bar_empty + 8u * s);
                        }
                        umma_commit_cg2_elect(bar_pempty + 8u * ps);      // both CTAs' copies of this patch have been read
                    }
                } else
                for (int kb0 = 0; kb0 < p.num_kb; kb0 += G) {
                    const uint32_t s = ms, ph = mph;
                    if (++ms == (uint32_t)S) { ms = 0; mph ^= 1u; }
                    if (pend && mbar_test(bar_a2full, q_ph)) issue_m2();
                    mbar_wait(bar_full + 8u * s, ph);
                    tc_fence_after();
                    const int gcount = p.num_kb - kb0 < G ? p.num_kb - kb0 : G;
                    for (int kk = 0; kk < gcount; kk++) {
                        const uint32_t sa = base + (s * (uint32_t)G + (uint32_t)kk) * (uint32_t)STAGE_BYTES;
                        const uint64_t da = make_smem_desc<SWZ>(sa);
                        const uint64_t db = make_smem_desc<SWZ>(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++)
                            umma_f16_cg2_elect(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb0 | kk | k) ? 1u : 0u);
                    }
                    umma_commit_cg2_elect(bar_empty + 8u * s);
                }
                umma_commit_cg2_elect(bar_tfull + 8u * as);
                if (pend) { mbar_wait(bar_a2full, q_ph); issue_m2(); }      // the previous tile's second MMA goes out before anything waits on its stage
                if (p.q_on) { pend = true; pend_as = as; }
            }
            if (pend) { mbar_wait(bar_a2full, q_ph); issue_m2(); }
        }
    } else {
        const int q = warp & 3;
        const int set = (warp - 2) >> 2;
        constexpr int NSETS = NEPI / 4;
        constexpr int NCH = BN / 32;
        const int r = q * 32 + lane;
        uint32_t ti = 0, sit = 0;
        const bool has_res = p.res != nullptr;
        const bool gw64 = p.epi_gw == 64;
        const uint32_t slab_bytes = gw64 ? 2u * kSlabBytes : kSlabBytes;
        const uint32_t my_slabs = slabs + (uint32_t)(warp - 2) * 2u * slab_bytes;
        if (p.epi && has_res && cluster_id < p.num_tiles)
            res_prefetch(p, my_slabs, tile_m0(cluster_id) + q * 32, tile_n0(cluster_id) + 32 * set, lane, gw64);
        const bool chain = p.q_on != 0;
        for (int ct = cluster_id; ct < p.num_tiles; ct += nclusters, ti++) {
            const long long m0 = tile_m0(ct);
            const int n0 = tile_n0(ct);
            const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
            const long long pr = m0 + r;
            bool valid = pr < p.M_total;
            if (chain && has_res && ti > 0) {
                // chain fusion: the second epilogue uses the slabs after the last group of a tile, so the skip tile of the next
                // tile's first group cannot be prefetched across the tile boundary; it is requested here, before the wait on the MMAs
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                res_prefetch(p, my_slabs + (sit & 1u) * slab_bytes, m0 + q * 32, n0 + 32 * set, lane, gw64);
            }
            long long drow = 0;
            if (p.mode == 2) {
                int bimg, boh0, bow0;
                const bool real = box_tile(ct, bimg, boh0, bow0);
                const int th = r / p.TW, tw = r - th * p.TW;
                const int oh = boh0 + th, ow = bow0 + tw;
                valid = real && (r < p.TH * p.TW) && oh < p.OH && ow < p.OW;
                drow = ((long long)bimg * (p.OH + 2) + oh + 1) * (p.OW + 2) + ow + 1;
            } else {
                const unsigned up = (unsigned)(valid ? pr : 0);
                const int wp = (int)(up % (unsigned)p.Wp);
                const int hp = (int)((up / (unsigned)p.Wp) % (unsigned)p.Hp);
                valid = valid && hp >= 1 && hp <= p.Hp - 2 && wp >= 1 && wp <= p.Wp - 2;
            }
            mbar_wait(bar_tfull + 8u * as, aph);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)BN;
            uint32_t va[32];
            [[maybe_unused]] uint32_t vb[32];
            if (p.mode == 2) {
                // strided boxes: output rows are not consecutive in memory -> per-thread stores (conv_tc.cuh epilogue_chunk); NEPI == 4
                if constexpr (NEPI == 4) {
                    tmem_ld32_issue(tacc, va);
#pragma unroll 1
                    for (int c0 = 0; c0 < BN; c0 += 64) {
                        tmem_ld_wait(va);
                        tmem_ld32_issue(tacc + (uint32_t)(c0 + 32), vb);
                        if (valid && n0 + c0 < p.cout_store) epilogue_chunk<false>(p, va, sbias, sbias, n0 + c0, drow, 0, 0, 0);
                        __syncwarp();
                        tmem_ld_wait(vb);
                        if (c0 + 64 < BN) tmem_ld32_issue(tacc + (uint32_t)(c0 + 64), va);
                        if (valid && n0 + c0 + 32 < p.cout_store) epilogue_chunk<false>(p, vb, sbias, sbias, n0 + c0 + 32, drow, 0, 0, 0);
                        __syncwarp();
                    }
                    tc_fence_before();
                    if (lane == 0) mbar_arrive_leader(bar_tempty + 8u * as);
                }
                continue;
            }
            auto do_group = [&](const uint32_t (&v)[32], int k) {
                if (n0 + 32 * k >= p.cout_store) return;
                const int h = gw64 ? (k & 1) : 0;
                const bool last = !gw64 || h == 1 || n0 + 32 * (k + 1) >= p.cout_store;
                const uint32_t slab = my_slabs + (sit & 1u) * slab_bytes;
                if (has_res && h == 0) { cp_async_wait_all(); __syncwarp(); }
                epi_group(p, v, sbias + n0 + 32 * k, slab, lane, h, valid, has_res, gw64);
                if (chain) {
                    // the fp16 result chunk just written to the slab is also the second MMA's A operand: rows = this CTA's 128
                    // pixels, 64-channel K blocks in the SWIZZLE_128B K-major layout
                    const int c2 = n0 + 32 * k - p.q_col0;
                    if (c2 >= 0 && c2 < 64 * p.q_kb) {
                        const uint32_t blk = a2 + (uint32_t)(c2 >> 6) * 16384u + (uint32_t)r * 128u;
                        const int ch0 = (c2 & 63) >> 3;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            sts128(blk + (uint32_t)(((ch0 + j) ^ (r & 7)) << 4), lds128(slab_chunk_addr(slab, lane, 4 * h + j, gw64)));
                    }
                }
                if (!last) return;
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                if (has_res) {
                    int nk = gw64 ? k + 1 : k + NSETS, nct = ct;
                    if (nk >= NCH || n0 + 32 * nk >= p.cout_store) { nk = set; nct = chain ? p.num_tiles : ct + nclusters; }
                    if (nct < p.num_tiles)
                        res_prefetch(p, my_slabs + ((sit + 1u) & 1u) * slab_bytes, tile_m0(nct) + q * 32, tile_n0(nct) + 32 * nk, lane, gw64);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const int c0 = n0 + 32 * (k - h);
                    const bool second = p.split_col && c0 >= p.split_col;               // sibling fusion (conv_tc.cuh TcConvDesc::out2)
                    tma_store_2d(second ? &p.tmOut2 : &p.tmOut, slab, second ? c0 - p.split_col : c0, (int)(m0 + q * 32));
                    bulk_commit();
                }
                sit++;
            };
            auto release_now = [&]() { tc_fence_before(); if (lane == 0) mbar_arrive_leader(bar_tempty + 8u * as); };
            auto release_acc = [&]() { if (!chain) release_now(); };      // chain fusion: the stage is handed back after the second epilogue
            if constexpr (NEPI >= 8) {
#pragma unroll 1
                for (int k = set; k < NCH; k += NSETS) {
                    tmem_ld32_issue(tacc + (uint32_t)(32 * k), va);
                    tmem_ld_wait(va);
                    if (k + NSETS >= NCH) release_acc();
                    do_group(va, k);
                    __syncwarp();
                }
            } else {
                tmem_ld32_issue(tacc + (uint32_t)(32 * set), va);
#pragma unroll 1
                for (int k = set; k < NCH; k += 2 * NSETS) {
                    const int k2 = k + NSETS;
                    tmem_ld_wait(va);
                    if (k2 < NCH) tmem_ld32_issue(tacc + (uint32_t)(32 * k2), vb); else release_acc();
                    do_group(va, k);
                    __syncwarp();
                    if (k2 < NCH) {
                        tmem_ld_wait(vb);
                        if (k2 + NSETS < NCH) tmem_ld32_issue(tacc + (uint32_t)(32 * (k2 + NSETS)), va); else release_acc();
                        do_group(vb, k2);
                        __syncwarp();
                    }
                }
            }
            if (chain) {
                // A2 complete for this warp's groups: visible to the tensor core, then tell the leader's MMA warp
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(bar_a2full);
                // ---- second epilogue: Q = 1x1 conv on the tile, accumulated into the same TMEM stage
                mbar_wait(bar_t2full, ti & 1u);
                tc_fence_after();
                const int nch2 = p.q_n / 32;
#pragma unroll 1
                for (int k2 = set; k2 < nch2; k2 += NSETS) {
                    tmem_ld32_issue(tacc + (uint32_t)(32 * k2), va);
                    tmem_ld_wait(va);
                    if (k2 + NSETS >= nch2) release_now();
                    if (32 * k2 < p.q_cout_store) {
                        const int h = gw64 ? (k2 & 1) : 0;
                        const bool last = !gw64 || h == 1 || 32 * (k2 + 1) >= p.q_cout_store;
                        const uint32_t slab = my_slabs + (sit & 1u) * slab_bytes;
                        epi_group(p, va, sbias2 + 32 * k2, slab, lane, h, valid, false, gw64, p.q_act);
                        if (last) {
                            if (lane == 0) bulk_wait_read<0>();                      // as do_group: the other slab's store has drained
                            __syncwarp();
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0) { tma_store_2d(&p.tmOutQ, slab, 32 * (k2 - h), (int)(m0 + q * 32)); bulk_commit(); }
                            sit++;
                        }
                    }
                    __syncwarp();
                }
                if (set >= nch2) release_now();               // a warp set without a group of Q still owes its arrival
            }
        }
        if (p.epi && lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                     // neither CTA may exit (or free TMEM) while the pair still uses its smem / barriers
    if (warp == 1) { tc_fence_after(); tmem_dealloc_cg2(tmem_base, TMEM_COLS); }
}

template <int BN, int NEPI>
inline cudaError_t launch_tc2(const TcConvPlan& pl, dim3 grid, cudaStream_t st) {
    static DeviceOnce once;
    if (once.first_use()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, NEPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024));
        if (e != cudaSuccess) { once.forget(); return e; }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(64 + 32 * NEPI); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BN, NEPI>, pl.p);
}

// launch of a cta2 plan (tc_plan2): one cluster per SM pair
inline int tc2_launch(const TcConvPlan& pl_in, int batch, cudaStream_t st) {
    TcConvPlan pl = pl_in;
    long long m_tiles;
    if (pl.kind == 2) {
        pl.p.M_total = (long long)batch * pl.p.tiles_per_img;   // box mode: number of 128-row M tiles (two per cluster tile)
        m_tiles = (pl.p.M_total + 1) / 2;
    } else {
        pl.p.M_total = (long long)batch * pl.p.Hp * pl.p.Wp;
        pl.p.m_start = pl.p.Wp + 1;                             // tiles start at the first interior pixel (see TcParams)
        m_tiles = (pl.p.M_total - 2 * pl.p.m_start + 255) / 256;
    }
    pl.p.num_tiles = (int)(m_tiles * pl.p.n_tiles);
    const int max_clusters = sm_count() / 2;
    const int nclusters = pl.p.num_tiles < max_clusters ? pl.p.num_tiles : max_clusters;
    dim3 grid((unsigned)(2 * nclusters));
    cudaError_t e = cudaErrorInvalidValue;
    if (pl.nepi == 16) {
        switch (pl.tile_n) {
            case 128: e = launch_tc2<128, 16>(pl, grid, st); break;
            case 256: e = launch_tc2<256, 16>(pl, grid, st); break;
        }
    } else if (pl.nepi == 8) {
        switch (pl.tile_n) {
            case 64: e = launch_tc2<64, 8>(pl, grid, st); break;
            case 128: e = launch_tc2<128, 8>(pl, grid, st); break;
            case 256: e = launch_tc2<256, 8>(pl, grid, st); break;
        }
    } else {
        switch (pl.tile_n) {
            case 64: e = launch_tc2<64, 4>(pl, grid, st); break;
            case 128: e = launch_tc2<128, 4>(pl, grid, st); break;
            case 256: e = launch_tc2<256, 4>(pl, grid, st); break;
        }
    }
    return e == cudaSuccess ? 0 : -1;
}

// Plan for the CTA-pair kernel; returns the kernel kind (1 flat, 2 strided box) when the layer is eligible, 0 otherwise.
inline int tc_plan2(const TcConvDesc& d, TcConvPlan* pl, std::string* err, int bn, int smem_budget_kb, int group, int nepi, int gw, int patch = 0) {
    if (d.pairx && (d.cin != 32 || d.stride != 2 || d.k != 3 || !d.w16_pair || d.in_ld != 32 || d.in_choff != 0)) return 0;
    const int cin = d.pairx ? 64 : d.cin;
    const int ntaps = d.pairx ? 6 : d.k * d.k;
    if (d.raw_in || cin % 64 != 0 || d.split || d.out_f32 || d.upsample) return 0;
    const bool box = d.stride == 2;
    if ((d.out2 || d.q_on) && box) return 0;
    if (box ? (d.k != 3 || nepi != 4 || gw != 32) : (d.stride != 1)) return 0;
    if (d.cout_pad % bn || (nepi != 4 && nepi != 8 && nepi != 16)) return 0;
    if (!box && (d.cout % gw != 0 || (gw != 32 && (gw != 64 || nepi != 4)))) return 0;
    // sixteen epilogue warps (four per TMEM lane quarter, one 32-column group each at N = 128, two at N = 256): four warps per
    // scheduler instead of two hide the TMEM-load / MUFU / shared-store latencies of the epilogue, which is what bounds the 1x1 layers
    if (nepi == 16 && (box || bn < 128 || d.q_on)) return 0;
    TcConvPlan P;
    TcParams& p = P.p;
    memset(&p, 0, sizeof(p));
    P.tile_n = bn; P.bk = 64; P.kind = box ? 2 : 1; P.nepi = nepi; P.cta2 = 1; P.ctas_per_sm = 1;
    p.mode = box ? 2 : 1; p.epi = box ? 0 : 1; p.epi_gw = gw;
    p.bias = d.bias; p.out = d.out; p.res = reinterpret_cast<const __half*>(d.res);
    p.act_scale = 1.f; p.inv_act_scale = 1.f;
    p.out_ld = d.out_ld; p.out_choff = d.out_choff; p.res_ld = d.res_ld; p.res_choff = d.res_choff;
    p.act = d.act;
    if (const char* env = getenv("Y4_DEBUG_ACT")) p.act = atoi(env);
    if (p.act == 2 && getenv("Y4_MISH_OLD")) p.act = 3;                     // A/B timing of the mish epilogue (act32_fast<3>)
    if (getenv("Y4_DEBUG_NORES")) p.res = nullptr;                          // timing experiments only (wrong results)
    p.cout_store = d.cout;
    p.ksize = d.k;
    p.kb_per_tap = cin / 64;
    p.num_kb = ntaps * p.kb_per_tap;
    p.pairx = d.pairx;
    const int K = ntaps * cin;
    const int in_Hp = d.in_H + 2, in_Wp = d.in_H + 2;
    P.in_Hp = in_Hp; P.in_Wp = in_Wp;
    p.Hp = in_Hp; p.Wp = in_Wp;
    p.rows_alloc = (long long)d.max_batch * in_Hp * in_Wp;
    char* in_base = reinterpret_cast<char*>(const_cast<void*>(d.in)) + (size_t)d.in_choff * 2;
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)d.cout_pad};
        cuuint64_t str[1] = {(cuuint64_t)K * 2};
        cuuint32_t box2[2] = {64, (cuuint32_t)(bn / 2)};
        if (!encode_map(&p.tmW, const_cast<__half*>(d.pairx ? d.w16_pair : d.w16), 2, dims, str, box2, 128, err)) return -1;
    }
    if (!box) {
        cuuint64_t dims[2] = {(cuuint64_t)d.cin, (cuuint64_t)p.rows_alloc};
        cuuint64_t str[1] = {(cuuint64_t)d.in_ld * 2};
        cuuint32_t box2[2] = {64, 128};
        if (!encode_map(&p.tmA[0], in_base, 2, dims, str, box2, 128, err)) return -1;
        cuuint64_t odims[2] = {(cuuint64_t)(d.out2 ? d.split_col : p.cout_store), (cuuint64_t)p.rows_alloc};
        cuuint64_t ostr[1] = {(cuuint64_t)d.out_ld * 2};
        cuuint32_t obox[2] = {(cuuint32_t)gw, 32};
        char* out_base = reinterpret_cast<char*>(d.out) + (size_t)d.out_choff * 2;
        if (!encode_map(&p.tmOut, out_base, 2, odims, ostr, obox, gw * 2, err)) return -1;
        if (d.out2) {
            if (d.res || d.split_col % gw != 0 || (d.cout - d.split_col) % gw != 0) return 0;
            cuuint64_t odims2[2] = {(cuuint64_t)(d.cout - d.split_col), (cuuint64_t)p.rows_alloc};
            cuuint64_t ostr2[1] = {(cuuint64_t)d.out2_ld * 2};
            char* out2_base = reinterpret_cast<char*>(d.out2) + (size_t)d.out2_choff * 2;
            if (!encode_map(&p.tmOut2, out2_base, 2, odims2, ostr2, obox, gw * 2, err)) return -1;
            p.split_col = d.split_col;
        }
    } else {
        // output box TH x TW per CTA (<= 128 pixels), chosen for the most useful rows per tile; four parity-plane views of the input
        p.OH = d.OH; p.OW = d.OH;
        int bestTW = 1, bestTH = 1; double best = -1;
        for (int tw = 1; tw <= d.OH && tw <= 128; tw++)
            for (int th = 1; th * tw <= 128 && th <= d.OH; th++) {
                const long long tiles = (long long)((d.OH + tw - 1) / tw) * ((d.OH + th - 1) / th);
                const double eff = (double)d.OH * d.OH / (tiles * 128.0);
                if (eff > best + 1e-9) { best = eff; bestTW = tw; bestTH = th; }
            }
        p.TW = bestTW; p.TH = bestTH;
        p.tiles_w = (d.OH + bestTW - 1) / bestTW;
        p.tiles_per_img = p.tiles_w * ((d.OH + bestTH - 1) / bestTH);
        for (int ph = 0; ph < 2; ph++)
            for (int pw = 0; pw < 2; pw++) {
                char* b = in_base + ((size_t)ph * in_Wp + pw) * d.in_ld * 2;
                cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)in_Wp / 2, (cuuint64_t)in_Hp / 2, (cuuint64_t)d.max_batch};
                cuuint64_t str[3] = {(cuuint64_t)2 * d.in_ld * 2, (cuuint64_t)2 * in_Wp * d.in_ld * 2, (cuuint64_t)in_Hp * in_Wp * d.in_ld * 2};
                cuuint32_t box4[4] = {64, (cuuint32_t)bestTW, (cuuint32_t)bestTH, 1};
                if (!encode_map(&p.tmA[ph * 2 + pw], b, 4, dims, str, box4, 128, err)) return -1;
            }
    }
    const bool deep_b = patch && group >= 2;                   // patch mode, `group` 2: as many B stages as fit beside three patches
    if (patch || group < 1) group = 1;
    if (group > p.num_kb) group = p.num_kb;
    p.group = group;
    const size_t epi_bytes = box ? 0 : epi_slab_bytes(nepi, gw);
    size_t stage_bytes = ((size_t)128 * 64 * 2 + (size_t)(bn / 2) * 64 * 2) * group;
    size_t chain_bytes = 0;
    if (d.q_on) {
        // chain fusion: one N tile must hold every channel Q reads; Q's N fits the accumulator stage; whole slab groups
        if (d.cout_pad != bn || d.q_cin % 64 != 0 || d.q_col0 % 64 != 0 || d.q_col0 + d.q_cin > d.cout || d.q_cout_pad > bn ||
            d.q_cout_pad % 32 != 0 || d.q_cout % gw != 0) return 0;
        p.q_on = 1; p.q_kb = d.q_cin / 64; p.q_n = d.q_cout_pad; p.q_col0 = d.q_col0; p.q_cout_store = d.q_cout; p.q_act = d.q_act;
        p.q_bias = d.q_bias;
        cuuint64_t wdims[2] = {(cuuint64_t)d.q_cin, (cuuint64_t)d.q_cout_pad};
        cuuint64_t wstr[1] = {(cuuint64_t)d.q_cin * 2};
        cuuint32_t wbox[2] = {64, (cuuint32_t)(d.q_cout_pad / 2)};
        if (!encode_map(&p.tmW2, const_cast<__half*>(d.q_w16), 2, wdims, wstr, wbox, 128, err)) return -1;
        cuuint64_t qdims[2] = {(cuuint64_t)d.q_cout, (cuuint64_t)p.rows_alloc};
        cuuint64_t qstr[1] = {(cuuint64_t)d.q_out_ld * 2};
        cuuint32_t qbox[2] = {(cuuint32_t)gw, 32};
        char* q_base = reinterpret_cast<char*>(d.q_out) + (size_t)d.q_out_choff * 2;
        if (!encode_map(&p.tmOutQ, q_base, 2, qdims, qstr, qbox, gw * 2, err)) return -1;
        chain_bytes = (size_t)p.q_kb * 16384 + (((size_t)p.q_kb * (size_t)(p.q_n / 2) * 128 + 1023) / 1024) * 1024;
    }
    const size_t fixed = 1024 + chain_bytes + epi_bytes + 16 * 8 + 224 + 4 * (size_t)d.cout_pad + 4 * (size_t)p.q_n;
    const size_t budget = (size_t)smem_budget_kb * 1024;
    size_t patch_total = 0;
    int S;
    if (patch) {
        // A-patch reuse (mode 3): 3x3 stride 1 only; one patch = 130 + 2*Wp rows of 128 B in 32-row TMA boxes; the ring holds B only
        if (box || d.k != 3 || patch != 1) return 0;
        p.mode = 3; p.patch_taps = 9; p.patch_box_rows = 32; p.patch_boxes = (130 + 2 * in_Wp + 31) / 32;
        p.patch_bytes = ((p.patch_boxes * 32 * 128 + 1023) / 1024) * 1024;
        cuuint64_t dims[2] = {(cuuint64_t)d.cin, (cuuint64_t)p.rows_alloc};
        cuuint64_t str[1] = {(cuuint64_t)d.in_ld * 2};
        cuuint32_t box2[2] = {64, 32};
        if (!encode_map(&p.tmA[1], in_base, 2, dims, str, box2, 128, err)) return -1;
        stage_bytes = (size_t)(bn / 2) * 64 * 2;
        S = 6;                                                         // B stages: a half B tile is small and re-fetched per tap
        if (deep_b) {
            // a B stage feeds four MMAs (256 tensor cycles at N = 128): six stages cover ~1,500 cycles of L2 latency, which is
            // borderline under load.  Deep variant: three patch slots (two if three do not fit), the rest of the budget as B stages.
            int best_s = 0, best_ps = 0;
            for (int ps = 3; ps >= 2 && !best_s; ps--) {
                if (budget < fixed + 64 + (size_t)ps * (size_t)p.patch_bytes) continue;
                int s2 = (int)((budget - fixed - 64 - (size_t)ps * (size_t)p.patch_bytes) / stage_bytes);
                if (s2 > 12) s2 = 12;
                if (s2 > 6) { best_s = s2; best_ps = ps; }
            }
            if (!best_s) return 0;                                     // nothing deeper than the default fits: no new plan
            S = best_s;
            p.patch_slots = best_ps;
        } else {
        const size_t bpart = (size_t)S * stage_bytes;
        if (budget < fixed + bpart + 2 * (size_t)p.patch_bytes) return 0;
        int PS = (int)((budget - fixed - bpart) / (size_t)p.patch_bytes);
        p.patch_slots = PS > 4 ? 4 : PS;
        }
        patch_total = (size_t)p.patch_slots * (size_t)p.patch_bytes;
    } else {
        S = budget > fixed ? (int)((budget - fixed) / stage_bytes) : 0;
        if (S > 8) S = 8;
        if (S < 2) return 0;
    }
    P.stages = S; p.stages = S;
    p.n_tiles = d.cout_pad / bn;
    p.bias_n = d.cout_pad;
    P.smem = 1024 + patch_total + S * stage_bytes + chain_bytes + epi_bytes + 16 * S + 224 + 4 * (size_t)p.bias_n + 4 * (size_t)p.q_n;
    if (P.smem > 225 * 1024) return 0;
    *pl = P;
    return P.kind;
}

}  // namespace y4
