// mnk_heads_mma.cu -- the policy / value head tails on tcgen05 (boards up to 96 cells).
//
// Same function as mnk_heads.cu (reference: src/alg/architectures/resnet.py:41-63),
//     policy: LayerNorm(2A) -> ReLU -> Linear(2A,128) -> LayerNorm(128) -> ReLU -> Linear(128,A)   = logits
//     value : LayerNorm(A)  -> ReLU -> Linear(A,128)  -> LayerNorm(128) -> ReLU -> Linear(128,1) -> Tanh
// but the three Linear layers with more than one output are UMMA GEMMs over a tile of 128 samples (M = 128 rows):
//     D1p[128 x 128] = LN1(policy_feat)[128 x K1p] * W1p^T     K1p = 2A rounded up to 16 (zero padded)
//     D1v[128 x 128] = LN1(value_feat) [128 x K1v] * W1v^T     K1v =  A rounded up to 16
//     D2 [128 x Np ] = LN2(D1p + b1)   [128 x 128] * W2^T      Np  =  A rounded up to 16
// with 16-bit operands (the tower's operand type, fp16 by default), fp32 accumulation in TMEM, LayerNorm / ReLU / bias
// in fp32 on the CUDA cores between the GEMMs, Linear(128,1) + Tanh of the value head as a dot product in the same pass.
// The fp32 kernel spent 0.20 ms per 32,768 samples (23 % FMA pipe, 42 % issue) on 2 % of the forward's FLOPs; a tile here
// is three short MMA chains (K1p/16 + K1v/16 + 8 instructions) and the kernel is bound by the 1.3 KB per sample it
// reads and writes.
//
// One persistent CTA per SM, 17 warps.  Per tile: (A) warps 0-15 normalise 8 rows each (coalesced loads, shuffle
// reductions) into the K-major no-swizzle operand layout [k-chunk][row][16 B]; (B) one thread issues the two first-layer
// GEMMs; (C) warp (head, lane quarter, column half) pulls 64 accumulator columns of its 32 rows into registers, the two
// halves exchange partial sums through shared memory, and the normalised row goes back to shared memory as the next
// operand (policy) or into the value dot product; (D) the logits GEMM; (E) + bias, staged through shared memory so that
// the tile's logits leave as one contiguous block.  All weights (130 KB at 9x9) stay in shared memory for the CTA's life.
#include <algorithm>

#include "mnk_dispatch.cuh"
#include "mnk_umma.cuh"

namespace hm {
using namespace mnk_umma;
constexpr int kH = 128;                 // head_hidden_dim of resnet_b_s
constexpr int kRows = 128;              // samples per tile = UMMA M
constexpr int kMaxCells = 96;            // shared memory: both first-layer operands + weights, 215 KB at 90 cells
constexpr int kEpiWarps = 16;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr int kTmemCols = 512;          // D1p at 0, D1v at 128, D2 at 256
constexpr float kEps = 1e-5f;           // torch.nn.LayerNorm default
constexpr int kHeadBarrier0 = 1;        // named barriers 1, 2: the 8 warps of a head exchange LayerNorm partial sums

struct Params {
    const float* pf;          // f32 [rows][2A]
    const float* vf;          // f32 [rows][A]   (unused when values == null)
    long long rows;
    int cells, k1p, k1v, np;  // A; 2A and A rounded up to 16; A rounded up to 16
    const unsigned char* w1p; // op16 [k1p/8][128][8]   element (k, n) = Linear(2A,128).weight[n][k]
    const unsigned char* w1v; // op16 [k1v/8][128][8]
    const unsigned char* w2;  // op16 [16][np][8]       element (k, n) = Linear(128,A).weight[n][k]
    const float* params;      // f32: p_ln1_w[2A] p_ln1_b[2A] v_ln1_w[A] v_ln1_b[A] p_b1 v_b1 p_ln2_w p_ln2_b v_ln2_w v_ln2_b v_w2 (128 each)
                              //      p_b2[A] v_b2[1]
    float* logits;            // f32 [rows][A]
    float* values;            // f32 [rows] or null
    int* error;
};

struct Layout {               // byte offsets into dynamic shared memory (all multiples of 128)
    int a1p, w1p, a1v, w1v, a2, w2, prm, part, bars, total;
};

__host__ __device__ inline Layout layout_for(int cells, int k1p, int k1v, int np) {
    Layout l;
    int off = 0;
    auto take = [&](int bytes) { const int at = off; off += (bytes + 127) & ~127; return at; };
    l.a1p = take(kRows * k1p * 2);          // also the fp32 logits staging [128][A] afterwards (A*4 <= k1p*2)
    l.w1p = take(kH * k1p * 2);
    l.a1v = take(kRows * k1v * 2);
    l.w1v = take(kH * k1v * 2);
    l.a2 = take(kRows * kH * 2);
    l.w2 = take(np * kH * 2);
    l.prm = take((6 * cells + 7 * kH + cells + 1) * 4);
    l.part = take(2 * kRows * 2 * 2 * 4);   // [head][row][column half][sum, sumsq]
    l.bars = take(64);
    l.total = off;
    return l;
}

MNK_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(MNK_FULL_WARP, v, o);
    return v;
}

MNK_DEV void tmem_ld32_issue(u32 taddr, u32 (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
MNK_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// LayerNorm + ReLU of one sample row (len values, lane-strided in registers) into the operand layout; one warp per row
template <int kItems>
MNK_DEV void ln1_row(const float* __restrict__ grow, bool live, int len, int kpad, const float* gamma, const float* beta,
                     unsigned char* a_op, int row, int lane) {
    float v[kItems];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int k = lane + 32 * i;
        v[i] = (k < len && live) ? __ldg(grow + k) : 0.f;
        sum += v[i];
    }
    const float mean = warp_sum(sum) / (float)len;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const float d = (lane + 32 * i < len) ? v[i] - mean : 0.f;
        sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)len + kEps);
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int k = lane + 32 * i;
        if (k < kpad) {
            const float y = (k < len) ? fmaxf((v[i] - mean) * rstd * gamma[k] + beta[k], 0.f) : 0.f;
            const u32 pk = act_pack2(y, 0.f);
            *reinterpret_cast<unsigned short*>(a_op + (size_t)(k >> 3) * (kRows * 16) + row * 16 + (k & 7) * 2) = (unsigned short)pk;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) heads_mma_kernel(Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = p.cells, two = 2 * p.cells;
    const Layout L = layout_for(p.cells, p.k1p, p.k1v, p.np);
    const bool want_value = p.values != nullptr;
    float* prm = reinterpret_cast<float*>(smem + L.prm);
    const float* p_ln1_w = prm;
    const float* p_ln1_b = prm + two;
    const float* v_ln1_w = prm + 2 * two;
    const float* v_ln1_b = prm + 2 * two + A;
    const float* vec = prm + 2 * two + 2 * A;      // 7 vectors of 128: p_b1 v_b1 p_ln2_w p_ln2_b v_ln2_w v_ln2_b v_w2
    const float* p_b2 = vec + 7 * kH;
    const float* v_b2 = p_b2 + A;
    float* part = reinterpret_cast<float*>(smem + L.part);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + L.bars);     // [0] GEMM1 done, [1] GEMM2 done
    unsigned int* tmem_slot = reinterpret_cast<unsigned int*>(smem + L.bars + 32);

    // ---- one-time setup: barriers, TMEM, weights and parameters into shared memory ----------------------------
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        const uint4* s1 = reinterpret_cast<const uint4*>(p.w1p);
        uint4* d1 = reinterpret_cast<uint4*>(smem + L.w1p);
        for (int i = tid; i < kH * p.k1p / 8; i += kThreads) d1[i] = __ldg(s1 + i);
        const uint4* s2 = reinterpret_cast<const uint4*>(p.w1v);
        uint4* d2 = reinterpret_cast<uint4*>(smem + L.w1v);
        if (want_value)
            for (int i = tid; i < kH * p.k1v / 8; i += kThreads) d2[i] = __ldg(s2 + i);
        const uint4* s3 = reinterpret_cast<const uint4*>(p.w2);
        uint4* d3 = reinterpret_cast<uint4*>(smem + L.w2);
        for (int i = tid; i < p.np * kH / 8; i += kThreads) d3[i] = __ldg(s3 + i);
        const int nprm = 6 * A + 7 * kH + A + 1;
        for (int i = tid; i < nprm; i += kThreads) prm[i] = __ldg(p.params + i);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = *tmem_slot;
    bool ok = true;
    const long long tiles = (p.rows + kRows - 1) / kRows;
    u32 phase = 0;

    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, phase ^= 1u) {
        const long long r0 = t * kRows;
        // ---- (A) LayerNorm(2A) / LayerNorm(A) + ReLU of 8 rows per warp into the operands --------------------
        if (warp < kEpiWarps) {
            for (int i = 0; i < kRows / kEpiWarps; ++i) {
                const int row = warp * (kRows / kEpiWarps) + i;
                const long long r = r0 + row;
                const bool live = r < p.rows;
                const long long rr = live ? r : 0;
                ln1_row<7>(p.pf + (size_t)rr * two, live, two, p.k1p, p_ln1_w, p_ln1_b, smem + L.a1p, row, lane);
                if (want_value) ln1_row<4>(p.vf + (size_t)rr * A, live, A, p.k1v, v_ln1_w, v_ln1_b, smem + L.a1v, row, lane);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- (B) first Linear of both heads ------------------------------------------------------------------
        if (warp == kMmaWarp) {
            if (elect_one()) {
                const u64 a_d = umma_desc(smem_u32(smem + L.a1p), kRows * 16, 128);
                const u64 b_d = umma_desc(smem_u32(smem + L.w1p), kH * 16, 128);
                for (int s = 0; s < p.k1p / 16; ++s)      // one MMA = two k-chunks = 2 * 2048 bytes further in both operands
                    umma_bf16(tmem, a_d + (u64)(s * 2 * kRows), b_d + (u64)(s * 2 * kH), umma_idesc_bf16(kH), s != 0);
                if (want_value) {
                    const u64 av_d = umma_desc(smem_u32(smem + L.a1v), kRows * 16, 128);
                    const u64 bv_d = umma_desc(smem_u32(smem + L.w1v), kH * 16, 128);
                    for (int s = 0; s < p.k1v / 16; ++s)
                        umma_bf16(tmem + kH, av_d + (u64)(s * 2 * kRows), bv_d + (u64)(s * 2 * kH), umma_idesc_bf16(kH), s != 0);
                }
                umma_commit(&bars[0]);
            }
            __syncwarp();
        }
        ok = __all_sync(MNK_FULL_WARP, mbar_wait(&bars[0], phase)) != 0 && ok;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- (C) + bias, LayerNorm(128), ReLU: warp = (head, lane quarter, column half) ---------------------
        if (warp < kEpiWarps) {
            const int head = warp >> 3, quarter = warp & 3, half = (warp >> 2) & 1;
            const int row = quarter * 32 + lane;
            if (head == 0 || want_value) {
                const float* b1 = vec + head * kH + 64 * half;
                const float* g2 = vec + (2 + 2 * head) * kH + 64 * half;
                const float* be2 = vec + (3 + 2 * head) * kH + 64 * half;
                u32 q0[32], q1[32];
                const u32 taddr = tmem + ((u32)(quarter * 32) << 16) + (u32)(head * kH + 64 * half);
                tmem_ld32_issue(taddr, q0);
                tmem_ld32_issue(taddr + 32, q1);
                tmem_wait_ld();
                float sum = 0.f, sq = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float x0 = __uint_as_float(q0[c]) + b1[c], x1 = __uint_as_float(q1[c]) + b1[32 + c];
                    q0[c] = __float_as_uint(x0);
                    q1[c] = __float_as_uint(x1);
                    sum += x0 + x1;
                    sq = fmaf(x0, x0, fmaf(x1, x1, sq));
                }
                float* mine = part + ((head * kRows + row) * 2 + half) * 2;
                mine[0] = sum;
                mine[1] = sq;
                asm volatile("bar.sync %0, %1;" ::"r"(kHeadBarrier0 + head), "r"(256) : "memory");
                const float* both = part + (head * kRows + row) * 4;
                const float mean = (both[0] + both[2]) * (1.0f / kH);
                const float var = fmaxf((both[1] + both[3]) * (1.0f / kH) - mean * mean, 0.f);
                const float rstd = rsqrtf(var + kEps);
                if (head == 0) {
                    unsigned char* a2 = smem + L.a2 + (size_t)(8 * half) * (kRows * 16) + row * 16;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {          // 8 k-chunks of 8 columns per 64-column half
                        u32 w[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const int c = 8 * ch + 2 * h;
                            const float x0 = __uint_as_float(c < 32 ? q0[c & 31] : q1[c & 31]);
                            const float x1 = __uint_as_float(c + 1 < 32 ? q0[(c + 1) & 31] : q1[(c + 1) & 31]);
                            const float y0 = fmaxf((x0 - mean) * rstd * g2[c] + be2[c], 0.f);
                            const float y1 = fmaxf((x1 - mean) * rstd * g2[c + 1] + be2[c + 1], 0.f);
                            w[h] = act_pack2(y0, y1);
                        }
                        *reinterpret_cast<uint4*>(a2 + (size_t)ch * (kRows * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                } else {                                          // Linear(128, 1) as a dot product, then Tanh
                    const float* w2v = vec + 6 * kH + 64 * half;
                    float dot = 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float y0 = fmaxf((__uint_as_float(q0[c]) - mean) * rstd * g2[c] + be2[c], 0.f);
                        const float y1 = fmaxf((__uint_as_float(q1[c]) - mean) * rstd * g2[32 + c] + be2[32 + c], 0.f);
                        dot = fmaf(y0, w2v[c], fmaf(y1, w2v[32 + c], dot));
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(kHeadBarrier0 + head), "r"(256) : "memory");    // partial sums consumed
                    mine[0] = dot;
                    asm volatile("bar.sync %0, %1;" ::"r"(kHeadBarrier0 + head), "r"(256) : "memory");
                    if (half == 0 && r0 + row < p.rows) p.values[r0 + row] = tanhf(both[0] + both[2] + v_b2[0]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- (D) logits GEMM ----------------------------------------------------------------------------------
        if (warp == kMmaWarp) {
            if (elect_one()) {
                const u64 a_d = umma_desc(smem_u32(smem + L.a2), kRows * 16, 128);
                const u64 b_d = umma_desc(smem_u32(smem + L.w2), p.np * 16, 128);
                const u32 idesc = umma_idesc_bf16(p.np);
                for (int s = 0; s < kH / 16; ++s)
                    umma_bf16(tmem + 2 * kH, a_d + (u64)(s * 2 * kRows), b_d + (u64)(s * 2 * p.np), idesc, s != 0);
                umma_commit(&bars[1]);
            }
            __syncwarp();
        }
        ok = __all_sync(MNK_FULL_WARP, mbar_wait(&bars[1], phase)) != 0 && ok;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- (E) + bias -> staging [128][A] f32 (over the dead policy operand) -> one contiguous block of logits --
        float* stage = reinterpret_cast<float*>(smem + L.a1p);
        if (warp < kEpiWarps) {
            const int quarter = warp & 3, cg = warp >> 2;      // column group of 32 (4 groups cover np <= 112)
            const int row = quarter * 32 + lane;
            if (32 * cg < p.np) {
                u32 q[32];
                tmem_ld32_issue(tmem + ((u32)(quarter * 32) << 16) + (u32)(2 * kH + 32 * cg), q);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int col = 32 * cg + c;
                    if (col < A) stage[row * A + col] = __uint_as_float(q[c]) + p_b2[col];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const long long live_rows = min((long long)kRows, p.rows - r0);
            const int count = (int)live_rows * A;
            float* dst = p.logits + (size_t)r0 * A;
            for (int i = tid; i < count; i += kThreads) dst[i] = stage[i];
        }
        __syncthreads();          // the staging area is the next tile's policy operand
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && p.error != nullptr) atomicMax(p.error, 2);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}
}  // namespace hm

extern "C" int mnk_resnet_heads_mma(const float* policy_feat, const float* value_feat, int64_t rows, int32_t cells, const void* w1p,
                                    const void* w1v, const void* w2, const float* params, float* logits, float* values,
                                    int32_t* error, void* stream) {
    if (!policy_feat || !w1p || !w1v || !w2 || !params || !logits || (values && !value_feat)) return MNK_ERR_NULL;
    if (rows < 0 || cells < 1) return MNK_ERR_ARG;
    if (cells > hm::kMaxCells) return MNK_ERR_GEOM;
    if ((reinterpret_cast<uintptr_t>(w1p) | reinterpret_cast<uintptr_t>(w1v) | reinterpret_cast<uintptr_t>(w2)) & 15u) return MNK_ERR_ALIGN;
    if (rows == 0) return MNK_OK;
    hm::Params p;
    p.pf = policy_feat; p.vf = value_feat; p.rows = rows; p.cells = cells;
    p.k1p = (2 * cells + 15) & ~15; p.k1v = (cells + 15) & ~15; p.np = (cells + 15) & ~15;
    p.w1p = static_cast<const unsigned char*>(w1p); p.w1v = static_cast<const unsigned char*>(w1v);
    p.w2 = static_cast<const unsigned char*>(w2); p.params = params; p.logits = logits; p.values = values; p.error = error;
    const hm::Layout lay = hm::layout_for(cells, p.k1p, p.k1v, p.np);
    const size_t smem = (size_t)lay.total;
    if (smem > 227 * 1024) return MNK_ERR_GEOM;
    static std::atomic<size_t> granted[kMaxDevices];
    if (int rc = mnk_optin_smem(hm::heads_mma_kernel, smem, granted)) return rc;
    const long long tiles = (rows + hm::kRows - 1) / hm::kRows;
    const unsigned grid = (unsigned)std::min<long long>(tiles, mnk_sm_count());
    hm::heads_mma_kernel<<<grid, hm::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return mnk_launch_status();
}
