// policy_forward.cu — the actor network of the rollout loop (BASELINE cfg 4) as ONE sm_100a kernel.
//
// The reference's actor is MultiObsEmbedding(ACTOR_CONFIGS) (src/model/network.py:34-196, src/model/attention.py:16-92,
// src/configs.py:134-153) with lidar / target / action-mask inputs: three 2-layer tanh embeddings -> 3 tokens x 128 ->
// one pre-norm transformer block (8 heads x 32, feed-forward 128) -> Linear(384, 128) tanh Linear(128, 2) tanh.  That is
// 1.23 MFLOP per env in 12 small GEMMs; run as ~25 PyTorch kernels the 65 536-env forward is bound by the activations it
// writes and re-reads (the (N, 3, 768) qkv tensor alone is 300 MB) and by 524 288 batched 3 x 3 attention products.
//
// Here a CTA of 8 warps owns 32 envs (96 token rows) and carries them through the whole network with every activation in
// shared memory: bf16 operands, float32 accumulation on the tensor cores (mma.sync m16n8k16; the GEMMs are M = 96 per CTA,
// far too small for a tcgen05 / TMEM pipeline to pay), float32 residual stream, LayerNorm, softmax and tanh in float32.
// Weights (bf16, fragment-packed by hope_policy_pack_matrix, K padded to 16) are read straight from L2 into B fragments, 256
// contiguous bytes per warp and fragment: each warp loads only the columns it owns, once per CTA.  Attention is folded into the head loop: per head, qkv for that head (N = 96) -> 3 x 3
// softmax per env -> the head's slice of to_out accumulated in registers, so the 768-wide qkv row never exists.
// HBM traffic per env: 668 B of float32 inputs in, 8 B out.  (The library is built with -fmad=false for the float64 env kernels, which
// must round every product and sum separately; the float32 arithmetic here has no such constraint and fuses explicitly with fmaf.)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/hope_b200.h"

namespace hope_policy {

constexpr int E = 128;                 // embedding width
constexpr int HEADS = 8, DH = 32;
constexpr int THREADS = 256, WARPS = THREADS / 32;
constexpr int KL = 128, KT = 16, KA = 48;             // padded input widths (120, 5, 42)
constexpr int LDX = E + 8;             // bf16 row stride of a 128-wide operand (272 B: ldmatrix rows fall on different banks)
constexpr int LDQ = 3 * DH + 8;        // qkv of one head
constexpr int LDA = DH + 8;            // attention output of one head
constexpr int LDR = E + 8;             // float32 residual stream (row stride = 8 banks mod 32: the float2 epilogue stores of 4 rows per phase do not collide)

// NM = 3: lidar, target, action mask (ACTOR_CONFIGS without the image).  NM = 4: the same plus the image token, whose encoder
// (conv stack -> Linear(2048, 256) tanh -> mean head) runs before this kernel; its tanh + re-embedding Linear(128, 128) is the
// "second embedding layer" of token 3 here.  Shared memory per CTA is NM x BM token rows: 32 envs for NM = 3 (106 KB, 2 CTAs
// per SM), 16 envs for NM = 4 (71 KB, 3 CTAs per SM).
template <int NM> struct Cfg {
    static constexpr int BM = NM == 3 ? 32 : 16;   // envs per CTA
    static constexpr int ROWS = NM * BM;           // token rows per CTA, row = modality * BM + env
    static constexpr int MTM = BM / 16;            // 16-row tiles of one modality
    static constexpr int MTA = ROWS / 16;          // 16-row tiles of all tokens
    static constexpr int MINB = NM == 3 ? 2 : 3;   // CTAs per SM
    static_assert(ROWS % (4 * WARPS) == 0 && MTA % 2 == 0 && ROWS <= 16 * WARPS, "LayerNorm passes, qkv tiling, one softmax row per lane pair");
};

template <int NM> struct Smem {
    static constexpr int BM = Cfg<NM>::BM, ROWS = Cfg<NM>::ROWS;
    float x[ROWS][LDR];                        // residual stream
    __nv_bfloat16 h[ROWS][LDX];                // embed hidden -> LayerNorm output -> bf16 copy of x for the output head
    union {
        struct { __nv_bfloat16 lidar[BM][KL + 8], target[BM][KT + 8], mask[BM][KA + 8]; } in;
        struct { __nv_bfloat16 qkv[ROWS][LDQ], att[ROWS][LDA]; } hd;
        __nv_bfloat16 ff[ROWS][LDX];           // feed-forward hidden
        float head[BM][E + 8];                 // hidden of the output head
    } u;
};

__device__ __forceinline__ float tanh_fast(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    // (not volatile: pure register in / out, so the compiler may interleave it with the fragment loads of the next K block)
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment of the 16 x 16 tile at (row0, k0) of a row-major bf16 matrix in shared memory
__device__ __forceinline__ void lda16x16(uint32_t (&a)[4], const __nv_bfloat16 *base, int ld, int row0, int k0, int lane) {
    const __nv_bfloat16 *p = base + (size_t)(row0 + (lane & 15)) * ld + k0 + ((lane >> 4) << 3);
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    // (reads shared memory: ordered against the stores and barriers around it by the memory clobber, free to move among the MMAs)
    asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr) : "memory");
}

// B fragments come from a FRAGMENT-PACKED copy of the weight W[n][k] (hope_policy_pack_matrix): for every 8-row tile nt and
// every 16-wide K block ks, the 32 lanes' register pairs (b0, b1) of the k16 x n8 operand lie side by side,
//   packed[((nt * KS + ks) * 32 + lane) * 4 + {0,1}] = W[8 nt + lane / 4][16 ks + 2 (lane % 4) + {0,1}]          (b0)
//   packed[((nt * KS + ks) * 32 + lane) * 4 + {2,3}] = W[8 nt + lane / 4][16 ks + 8 + 2 (lane % 4) + {0,1}]      (b1)
// so a warp reads one fragment as 256 contiguous bytes (2 L1 wavefronts).  Reading them from the row-major matrix instead
// (4 bytes per lane from 8 different rows, twice) costs 16 wavefronts per fragment and made the kernel LSU-bound: 85 % of the
// L1 data-pipe cycles, 62 % of them these loads (ncu, profiles/r02_ncu_full_summary_p.txt).
__device__ __forceinline__ void ldb16x8(uint32_t &b0, uint32_t &b1, const uint2 *frag /* packed + (nt * KS + ks) * 32 */, int lane) {
    const uint2 v = __ldg(frag + lane);
    b0 = v.x; b1 = v.y;
}

// acc[mi][ni] += A[row0 + 16 mi .. +15][0 .. K) * W[wrow(ni) .. +7][0 .. K)^T for MT row tiles and NT column tiles.
// a_row_of_k: row offset added per K block of E (the output head reads env e's three tokens as one 384-wide row).
template <int MT, int NT, int K, typename RowFn>
__device__ __forceinline__ void warp_gemm(float (&acc)[MT][NT][4], const __nv_bfloat16 *a, int lda, int row0, const uint2 *w, int ks_total, int kb0, RowFn wrow,
                                          int lane, int a_rows_per_kblock = 0) {
    const uint2 *wr[NT];  // this warp's column tiles: first fragment of each
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) wr[ni] = w + ((size_t)(wrow(ni) >> 3) * ks_total + kb0) * 32;
    // the weights come from L2 (300+ cycles): the B fragments run PF K blocks ahead of the tensor-core work in a register ring
    constexpr int KB = K / 16, PF = KB < 4 ? KB : 4;
    uint32_t b[PF][NT][2];
#pragma unroll
    for (int p = 0; p < PF; ++p)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) ldb16x8(b[p][ni][0], b[p][ni][1], wr[ni] + 32 * p, lane);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        const int k0 = 16 * kb, cur = kb % PF;
        const int arow = row0 + (k0 / E) * a_rows_per_kblock, ak = a_rows_per_kblock ? k0 % E : k0;
        uint32_t af[MT][4];  // all row tiles' A fragments first: their shared-memory latency is paid once per K block, not once per tile
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) lda16x16(af[mi], a, lda, arow + 16 * mi, ak, lane);
        uint32_t bc[NT][2];
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) { bc[ni][0] = b[cur][ni][0]; bc[ni][1] = b[cur][ni][1]; }
        if (kb + PF < KB) {  // refill the slot just consumed
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) ldb16x8(b[cur][ni][0], b[cur][ni][1], wr[ni] + 32 * (kb + PF), lane);
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) mma16816(acc[mi][ni], af[mi], bc[ni][0], bc[ni][1]);
    }
}

// accumulators start at the bias of their columns (column pair ccol, ccol + 1 of tile ni starting at col0 + 8 ni), so the epilogue has no add
template <int MT, int NT>
__device__ __forceinline__ void init_bias(float (&acc)[MT][NT][4], const float *__restrict__ bias, int col0, int ccol) {
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
        const float b0 = __ldg(bias + col0 + 8 * ni + ccol), b1 = __ldg(bias + col0 + 8 * ni + ccol + 1);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) { acc[mi][ni][0] = b0; acc[mi][ni][1] = b1; acc[mi][ni][2] = b0; acc[mi][ni][3] = b1; }
    }
}

template <int MT, int NT>
__device__ __forceinline__ void zero(float (&acc)[MT][NT][4]) {
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.f;
}

// LayerNorm (eps 1e-5, affine) of the residual rows -> bf16 operand.  8 lanes per row (16 columns each), 4 rows per warp pass:
// three shuffle steps per reduction instead of five, and four independent rows in flight per warp.
template <int NM>
__device__ __forceinline__ void layer_norm_rows(Smem<NM> &sm, const float *__restrict__ g, const float *__restrict__ b, int warp, int lane) {
    constexpr int ROWS = Cfg<NM>::ROWS;
    const int sub = lane >> 3, part = lane & 7;  // row within the pass; the lane's columns are 32 v + 4 part .. + 3, v = 0..3, so that the
                                                 // 8 lanes of a row read 128 contiguous bytes per load (no bank conflicts)
    float4 gg[4], bb[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) { gg[v] = __ldg(reinterpret_cast<const float4 *>(g) + 8 * v + part); bb[v] = __ldg(reinterpret_cast<const float4 *>(b) + 8 * v + part); }
#pragma unroll
    for (int pass = 0; pass < ROWS / (4 * WARPS); ++pass) {
        const int r = (pass * WARPS + warp) * 4 + sub;
        float4 x[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) x[v] = *reinterpret_cast<const float4 *>(&sm.x[r][32 * v + 4 * part]);
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) s += (x[v].x + x[v].y) + (x[v].z + x[v].w);
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s * (1.f / E);
        float q = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            x[v].x -= mean; x[v].y -= mean; x[v].z -= mean; x[v].w -= mean;
            q = fmaf(x[v].x, x[v].x, q); q = fmaf(x[v].y, x[v].y, q); q = fmaf(x[v].z, x[v].z, q); q = fmaf(x[v].w, x[v].w, q);
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float rstd = rsqrtf(q * (1.f / E) + 1e-5f);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf(x[v].x * rstd, gg[v].x, bb[v].x), fmaf(x[v].y * rstd, gg[v].y, bb[v].y));
            const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf(x[v].z * rstd, gg[v].z, bb[v].z), fmaf(x[v].w * rstd, gg[v].w, bb[v].w));
            *reinterpret_cast<uint2 *>(&sm.h[r][32 * v + 4 * part]) = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        }
    }
}

template <int NM>
__global__ void __launch_bounds__(THREADS, Cfg<NM>::MINB) k_policy_forward(int n, const float *__restrict__ lidar, const float *__restrict__ target, const float *__restrict__ mask,
                                                                        const float *__restrict__ img_mean, hope_policy_weights W, float *__restrict__ out) {
    constexpr int BM = Cfg<NM>::BM, ROWS = Cfg<NM>::ROWS, MTM = Cfg<NM>::MTM, MTA = Cfg<NM>::MTA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<NM> &sm = *reinterpret_cast<Smem<NM> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int env0 = blockIdx.x * BM;
    using fragp = const uint2 *;
    fragp w1_lidar = static_cast<fragp>(W.w1_lidar), w1_target = static_cast<fragp>(W.w1_target), w1_mask = static_cast<fragp>(W.w1_mask);
    fragp w_qkv = static_cast<fragp>(W.w_qkv), w_out = static_cast<fragp>(W.w_out), w_ff1 = static_cast<fragp>(W.w_ff1), w_ff2 = static_cast<fragp>(W.w_ff2),
          w_o1 = static_cast<fragp>(W.w_o1);
    const int crow = lane >> 2, ccol = (lane & 3) << 1;  // this lane's place in a 16 x 8 accumulator tile: rows crow, crow + 8; columns ccol, ccol + 1

    // ---- inputs -> bf16, zero padded --------------------------------------------------------------------------
    // (rows of lidar are 480 B, of the mask 168 B: both multiples of 8, so float2 loads stay aligned for every env)
    for (int i = tid; i < BM * (KL / 2); i += THREADS) {
        const int e = i / (KL / 2), c = 2 * (i % (KL / 2));
        float2 v = make_float2(0.f, 0.f);
        if (c < 120 && env0 + e < n) v = __ldg(reinterpret_cast<const float2 *>(lidar + (size_t)(env0 + e) * 120 + c));
        *reinterpret_cast<__nv_bfloat162 *>(&sm.u.in.lidar[e][c]) = __floats2bfloat162_rn(v.x, v.y);
    }
    for (int i = tid; i < BM * KT; i += THREADS) {
        const int e = i / KT, c = i % KT;
        sm.u.in.target[e][c] = __float2bfloat16((c < 5 && env0 + e < n) ? __ldg(target + (size_t)(env0 + e) * 5 + c) : 0.f);
    }
    for (int i = tid; i < BM * (KA / 2); i += THREADS) {
        const int e = i / (KA / 2), c = 2 * (i % (KA / 2));
        float2 v = make_float2(0.f, 0.f);
        if (c < 42 && env0 + e < n) v = __ldg(reinterpret_cast<const float2 *>(mask + (size_t)(env0 + e) * 42 + c));
        *reinterpret_cast<__nv_bfloat162 *>(&sm.u.in.mask[e][c]) = __floats2bfloat162_rn(v.x, v.y);
    }
    if constexpr (NM == 4) {  // re_embed_img.0: tanh of the image encoder's mean head = the hidden layer of token 3
        for (int i = tid; i < BM * (E / 2); i += THREADS) {
            const int e = i / (E / 2), c = 2 * (i % (E / 2));
            float2 v = make_float2(0.f, 0.f);
            if (env0 + e < n) v = __ldg(reinterpret_cast<const float2 *>(img_mean + (size_t)(env0 + e) * E + c));
            *reinterpret_cast<__nv_bfloat162 *>(&sm.h[3 * BM + e][c]) = __floats2bfloat162_rn(tanh_fast(v.x), tanh_fast(v.y));
        }
    }
    __syncthreads();

    // ---- embeddings, layer 1: h[m * BM + e] = tanh(in_m W1_m^T + b1_m).  Warp w owns columns 16 w .. 16 w + 15 of every modality ----
    const int ncol0 = 16 * warp;
    auto cols = [&](int ni) { return ncol0 + 8 * ni; };
    {
        float acc[MTM][2][4];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            init_bias(acc, W.b1[m], ncol0, ccol);
            if (m == 0) warp_gemm<MTM, 2, KL>(acc, &sm.u.in.lidar[0][0], KL + 8, 0, w1_lidar, KL / 16, 0, cols, lane);
            if (m == 1) warp_gemm<MTM, 2, KT>(acc, &sm.u.in.target[0][0], KT + 8, 0, w1_target, KT / 16, 0, cols, lane);
            if (m == 2) warp_gemm<MTM, 2, KA>(acc, &sm.u.in.mask[0][0], KA + 8, 0, w1_mask, KA / 16, 0, cols, lane);
#pragma unroll
            for (int mi = 0; mi < MTM; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const int c = ncol0 + 8 * ni + ccol, r = m * BM + 16 * mi + crow;
                    *reinterpret_cast<__nv_bfloat162 *>(&sm.h[r][c]) = __floats2bfloat162_rn(tanh_fast(acc[mi][ni][0]), tanh_fast(acc[mi][ni][1]));
                    *reinterpret_cast<__nv_bfloat162 *>(&sm.h[r + 8][c]) = __floats2bfloat162_rn(tanh_fast(acc[mi][ni][2]), tanh_fast(acc[mi][ni][3]));
                }
        }
    }
    __syncthreads();
    // ---- embeddings, layer 2: x = h W2_m^T + b2_m (float32 residual stream) --------------------------------------
    {
        float acc[MTM][2][4];
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            init_bias(acc, m < 3 ? W.b2[m < 3 ? m : 0] : W.b2_img, ncol0, ccol);
            warp_gemm<MTM, 2, E>(acc, &sm.h[0][0], LDX, m * BM, static_cast<fragp>(m < 3 ? W.w2[m < 3 ? m : 0] : W.w2_img), E / 16, 0, cols, lane);
#pragma unroll
            for (int mi = 0; mi < MTM; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const int c = ncol0 + 8 * ni + ccol, r = m * BM + 16 * mi + crow;
                    *reinterpret_cast<float2 *>(&sm.x[r][c]) = make_float2(acc[mi][ni][0], acc[mi][ni][1]);
                    *reinterpret_cast<float2 *>(&sm.x[r + 8][c]) = make_float2(acc[mi][ni][2], acc[mi][ni][3]);
                }
        }
    }
    __syncthreads();

    // ---- attention block: x += to_out(softmax(q k^T / sqrt(32)) v) over the 3 tokens of each env, pre-norm -----------
    layer_norm_rows(sm, W.ln1_g, W.ln1_b, warp, lane);
    __syncthreads();
    {
        float oacc[MTA][2][4];  // this warp's 16 columns of to_out for all token rows, accumulated over the heads
        init_bias(oacc, W.b_out, ncol0, ccol);
        // qkv of one head: ROWS rows x 96 columns = MTA x 12 tiles; warp w takes row tiles QM (w / 4) .. + QM - 1 and column tiles 3 (w % 4) .. + 2
        constexpr int QM = MTA / 2;
        const int qm0 = QM * (warp >> 2), qn0 = 3 * (warp & 3);
        for (int hd = 0; hd < HEADS; ++hd) {
            float qacc[QM][3][4];
            zero(qacc);
            auto qrow = [&](int ni) { const int t = qn0 + ni; return (t >> 2) * (HEADS * DH) + hd * DH + (t & 3) * 8; };  // q | k | v blocks of to_qkv, head hd
            warp_gemm<QM, 3, E>(qacc, &sm.h[0][0], LDX, 16 * qm0, w_qkv, E / 16, 0, qrow, lane);
#pragma unroll
            for (int mi = 0; mi < QM; ++mi)
#pragma unroll
                for (int ni = 0; ni < 3; ++ni) {
                    const int c = 8 * (qn0 + ni) + ccol, r = 16 * (qm0 + mi) + crow;
                    *reinterpret_cast<__nv_bfloat162 *>(&sm.u.hd.qkv[r][c]) = __floats2bfloat162_rn(qacc[mi][ni][0], qacc[mi][ni][1]);
                    *reinterpret_cast<__nv_bfloat162 *>(&sm.u.hd.qkv[r + 8][c]) = __floats2bfloat162_rn(qacc[mi][ni][2], qacc[mi][ni][3]);
                }
            __syncthreads();
            // softmax over the env's NM tokens: two threads per query row, each owns 16 of the head's 32 dims (its half of every q . k,
            // summed with one shuffle, and its half of the output)
            {
                // lanes L and L + 16 of a warp share query row 16 warp + L % 16 (first / second half of the head's dims): the 8 lanes
                // of a 128-bit shared-memory phase then sit in 8 different rows (row strides of 13 and 5 chunks of 16 B: no conflicts)
                const int r = (tid >> 5) * 16 + (lane & 15), half = lane >> 4, e = r % BM;
                const bool on = r < ROWS;
                float qv[16];
                float s[NM];
#pragma unroll
                for (int j = 0; j < NM; ++j) s[j] = 0.f;
                if (on) {
                    const uint4 *qp = reinterpret_cast<const uint4 *>(&sm.u.hd.qkv[r][16 * half]);
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const uint4 w4 = qp[v];
                        const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) { qv[8 * v + 2 * c] = __uint_as_float(ws[c] << 16); qv[8 * v + 2 * c + 1] = __uint_as_float(ws[c] & 0xffff0000u); }
                    }
#pragma unroll
                    for (int j = 0; j < NM; ++j) {
                        const uint4 *kp = reinterpret_cast<const uint4 *>(&sm.u.hd.qkv[j * BM + e][DH + 16 * half]);
                        float d[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains
#pragma unroll
                        for (int v = 0; v < 2; ++v) {
                            const uint4 w4 = kp[v];
                            const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) d[c] = fmaf(qv[8 * v + 2 * c + 1], __uint_as_float(ws[c] & 0xffff0000u), fmaf(qv[8 * v + 2 * c], __uint_as_float(ws[c] << 16), d[c]));
                        }
                        s[j] = (d[0] + d[1]) + (d[2] + d[3]);
                    }
                }
#pragma unroll
                for (int j = 0; j < NM; ++j) s[j] = (s[j] + __shfl_xor_sync(0xffffffffu, s[j], 16)) * 0.17677669529663687f;  // dim_head ** -0.5
                if (on) {
                    float mx = s[0];
#pragma unroll
                    for (int j = 1; j < NM; ++j) mx = fmaxf(mx, s[j]);
                    float pw[NM], den = 0.f;
#pragma unroll
                    for (int j = 0; j < NM; ++j) { pw[j] = __expf(s[j] - mx); den += pw[j]; }
                    const float inv = 1.f / den;
                    uint4 *dst = reinterpret_cast<uint4 *>(&sm.u.hd.att[r][16 * half]);
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < NM; ++j) {
                            const uint4 a4 = reinterpret_cast<const uint4 *>(&sm.u.hd.qkv[j * BM + e][2 * DH + 16 * half])[v];
                            const uint32_t as[4] = {a4.x, a4.y, a4.z, a4.w};
                            const float wj = pw[j] * inv;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                lo[c] = j == 0 ? wj * __uint_as_float(as[c] << 16) : fmaf(wj, __uint_as_float(as[c] << 16), lo[c]);
                                hi[c] = j == 0 ? wj * __uint_as_float(as[c] & 0xffff0000u) : fmaf(wj, __uint_as_float(as[c] & 0xffff0000u), hi[c]);
                            }
                        }
                        uint32_t o[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const __nv_bfloat162 pk = __floats2bfloat162_rn(lo[c], hi[c]);
                            o[c] = *reinterpret_cast<const uint32_t *>(&pk);
                        }
                        dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
            __syncthreads();
            // the head's slice of to_out: oacc += att (96 x 32) * w_out[:, 32 hd .. 32 hd + 31]^T
            warp_gemm<MTA, 2, DH>(oacc, &sm.u.hd.att[0][0], LDA, 0, w_out, HEADS * DH / 16, hd * (DH / 16), cols, lane);
        }
#pragma unroll
        for (int mi = 0; mi < MTA; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int c = ncol0 + 8 * ni + ccol, r = 16 * mi + crow;
                float2 *p = reinterpret_cast<float2 *>(&sm.x[r][c]), *q = reinterpret_cast<float2 *>(&sm.x[r + 8][c]);
                *p = make_float2(p->x + oacc[mi][ni][0], p->y + oacc[mi][ni][1]);
                *q = make_float2(q->x + oacc[mi][ni][2], q->y + oacc[mi][ni][3]);
            }
    }
    __syncthreads();

    // ---- feed-forward block: x += W2 tanh(W1 LN(x) + b1) + b2 ---------------------------------------------------------
    layer_norm_rows(sm, W.ln2_g, W.ln2_b, warp, lane);
    __syncthreads();
    {
        float acc[MTA][2][4];
        init_bias(acc, W.b_ff1, ncol0, ccol);
        warp_gemm<MTA, 2, E>(acc, &sm.h[0][0], LDX, 0, w_ff1, E / 16, 0, cols, lane);
#pragma unroll
        for (int mi = 0; mi < MTA; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int c = ncol0 + 8 * ni + ccol, r = 16 * mi + crow;
                *reinterpret_cast<__nv_bfloat162 *>(&sm.u.ff[r][c]) = __floats2bfloat162_rn(tanh_fast(acc[mi][ni][0]), tanh_fast(acc[mi][ni][1]));
                *reinterpret_cast<__nv_bfloat162 *>(&sm.u.ff[r + 8][c]) = __floats2bfloat162_rn(tanh_fast(acc[mi][ni][2]), tanh_fast(acc[mi][ni][3]));
            }
        __syncthreads();
        init_bias(acc, W.b_ff2, ncol0, ccol);
        warp_gemm<MTA, 2, E>(acc, &sm.u.ff[0][0], LDX, 0, w_ff2, E / 16, 0, cols, lane);
        // the residual stream's last use is the output head's bf16 operand: write x + ff straight into h
#pragma unroll
        for (int mi = 0; mi < MTA; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int c = ncol0 + 8 * ni + ccol, r = 16 * mi + crow;
                const float2 p = *reinterpret_cast<const float2 *>(&sm.x[r][c]), q = *reinterpret_cast<const float2 *>(&sm.x[r + 8][c]);
                *reinterpret_cast<__nv_bfloat162 *>(&sm.h[r][c]) = __floats2bfloat162_rn(p.x + acc[mi][ni][0], p.y + acc[mi][ni][1]);
                *reinterpret_cast<__nv_bfloat162 *>(&sm.h[r + 8][c]) = __floats2bfloat162_rn(q.x + acc[mi][ni][2], q.y + acc[mi][ni][3]);
            }
    }
    __syncthreads();  // h complete (all warps also passed their last read of u.ff before this point: u.head may be written now)

    // ---- output head: tanh(W_o2 tanh(W_o1 [x_0 | .. | x_NM-1] + b_o1) + b_o2) ---------------------------------------------
    {
        float acc[MTM][2][4];
        init_bias(acc, W.b_o1, ncol0, ccol);
        warp_gemm<MTM, 2, NM * E>(acc, &sm.h[0][0], LDX, 0, w_o1, NM * E / 16, 0, cols, lane, BM);  // K block m reads token rows m * BM + env
#pragma unroll
        for (int mi = 0; mi < MTM; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int c = ncol0 + 8 * ni + ccol, r = 16 * mi + crow;
                *reinterpret_cast<float2 *>(&sm.u.head[r][c]) = make_float2(tanh_fast(acc[mi][ni][0]), tanh_fast(acc[mi][ni][1]));
                *reinterpret_cast<float2 *>(&sm.u.head[r + 8][c]) = make_float2(tanh_fast(acc[mi][ni][2]), tanh_fast(acc[mi][ni][3]));
            }
    }
    __syncthreads();
    {   // Linear(128, 2): 4 lanes per (env, output), 32 products each
        const int pair = tid >> 2, part = tid & 3, e = (pair >> 1) % BM, o = pair & 1;  // 2 BM pairs x 4 lanes (whole warps either way)
        const float *hrow = &sm.u.head[e][32 * part], *wrow = W.w_o2 + o * E + 32 * part;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) s = fmaf(hrow[c], __ldg(wrow + c), s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0 && pair < 2 * BM && env0 + e < n) out[(size_t)(env0 + e) * 2 + o] = tanhf(s + __ldg(W.b_o2 + o));
    }
}

}  // namespace hope_policy

template <int NM>
static int launch_policy(int n, const float *d_lidar, const float *d_target, const float *d_mask, const float *d_img_mean, const hope_policy_weights *w, float *d_out,
                         void *stream) {
    using namespace hope_policy;
    const int smem = (int)sizeof(Smem<NM>);
    if (cudaFuncSetAttribute(k_policy_forward<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return HOPE_ERR_CUDA;
    constexpr int BM = Cfg<NM>::BM;
    k_policy_forward<NM><<<(n + BM - 1) / BM, THREADS, smem, static_cast<cudaStream_t>(stream)>>>(n, d_lidar, d_target, d_mask, d_img_mean, *w, d_out);
    return cudaGetLastError() == cudaSuccess ? HOPE_OK : HOPE_ERR_CUDA;
}

extern "C" {

int hope_policy_forward(int n, const float *d_lidar, const float *d_target, const float *d_mask, const hope_policy_weights *w, float *d_out, void *stream) {
    if (n <= 0 || !d_lidar || !d_target || !d_mask || !w || !d_out) return HOPE_ERR_INVALID;
    return launch_policy<3>(n, d_lidar, d_target, d_mask, nullptr, w, d_out, stream);
}

int hope_policy_forward_img(int n, const float *d_lidar, const float *d_target, const float *d_mask, const float *d_img_mean, const hope_policy_weights *w,
                            float *d_out, void *stream) {
    if (n <= 0 || !d_lidar || !d_target || !d_mask || !d_img_mean || !w || !d_out || !w->w2_img || !w->b2_img) return HOPE_ERR_INVALID;
    return launch_policy<4>(n, d_lidar, d_target, d_mask, d_img_mean, w, d_out, stream);
}

int hope_policy_forward_smem_bytes(void) { return (int)sizeof(hope_policy::Smem<3>); }

int hope_policy_pack_matrix(const float *h_w, int n_out, int n_in, int k_pad, void *h_packed) {
    if (!h_w || !h_packed || n_out <= 0 || n_in <= 0 || n_out % 8 || k_pad % 16 || k_pad < n_in) return HOPE_ERR_INVALID;
    uint16_t *out = static_cast<uint16_t *>(h_packed);
    const int ks_total = k_pad / 16;
    auto bf16 = [](float f) {  // round to nearest even, like torch's .to(bfloat16)
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN stays NaN
        u += 0x7fffu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    for (int nt = 0; nt < n_out / 8; ++nt)
        for (int ks = 0; ks < ks_total; ++ks)
            for (int lane = 0; lane < 32; ++lane)
                for (int half = 0; half < 2; ++half)
                    for (int e = 0; e < 2; ++e) {
                        const int row = 8 * nt + lane / 4, col = 16 * ks + 8 * half + 2 * (lane % 4) + e;
                        out[(((size_t)nt * ks_total + ks) * 32 + lane) * 4 + 2 * half + e] = col < n_in ? bf16(h_w[(size_t)row * n_in + col]) : (uint16_t)0;
                    }
    return HOPE_OK;
}

}  // extern "C"
