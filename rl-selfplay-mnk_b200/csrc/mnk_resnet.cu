// mnk_resnet.cu -- the residual tower of the reference's default policy/value network on tcgen05.
//
// Network (reference: src/alg/architectures/resnet.py:8-95 with the "resnet_b_s" parameters of
// src/alg/architectures/configs.py:28-35): conv3x3(2->32)+BN+ReLU, then `blocks` x
// [conv3x3+BN+ReLU, conv3x3+BN, +skip, ReLU], then the 1x1 convolutions that open the policy head
// (32->2) and the value head (32->1).  Eval-mode BatchNorm is folded into the convolution weights
// and biases on the host.  These convolutions are 98% of the forward's FLOPs and the
// only dense contraction on the hot path; the heads' LayerNorm/Linear layers that follow operate on
// the [N, 2*cells] / [N, cells] features this kernel writes.
//
// Formulation: implicit GEMM, D[pixel][c_out] = sum_{tap, c_in} A[pixel + off(tap)][c_in] * W[tap][c_out][c_in].
//   * A CTA owns SPC consecutive envs.  Each env's m x n board is laid out on pixel rows with row
//     stride PW = n+1 (the same guard-column trick as the bitboards) and PW+1 zero rows between
//     envs, so a 3x3 tap is a constant row offset off = dr*PW + dc and out-of-board neighbours read
//     zeros.  The tile is 1024 pixel rows = 8 UMMA M-blocks of 128.
//   * Activations live in shared memory as bf16 in the canonical K-major NO-SWIZZLE UMMA layout
//     [k-chunk (8 channels = 16 B)][row][16 B] with SBO = 128 B, i.e. consecutive rows are 16 B
//     apart: shifting the operand by `off` rows is just a different start address in the shared
//     memory descriptor.  No im2col, no data movement between layers: the whole tower runs out of
//     two ping-pong activation buffers (A0 scratch, A1 running feature map).
//   * One thread issues tcgen05.mma (M=128, N=32, K=16, bf16 -> fp32): 18 per M-block per layer,
//     accumulating in TMEM (8 blocks x 32 columns = 256 columns), and commits each block to an
//     mbarrier.  Four epilogue warps (one per TMEM lane quarter) wait per block, tcgen05.ld the 32
//     fp32 channels of their pixel row, add bias (+ skip), ReLU, write bf16 back to shared memory --
//     so the epilogue of block j overlaps the MMAs of blocks j+1...  Guard rows are rewritten as zeros.
//   * Layer weights (18 KB) stream through a 2-deep shared-memory ring with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx) issued one layer ahead.
//   * The input is read straight from the packed bitboards (72 B per env at 9x9) and the last
//     epilogue applies the two 1x1 head convolutions from fp32 registers and stores the features.
#include "mnk_dispatch.cuh"
#include "mnk_umma.cuh"


namespace rn {
constexpr int kC = 32;                        // tower channels
constexpr int kBlocksM = 4;                   // UMMA M-blocks per CTA (two CTAs per SM: one's epilogue overlaps the other's MMAs)
constexpr int kRows = 128 * kBlocksM;         // pixel rows per CTA tile
constexpr int kMargin = 24;                   // zero rows before/after the tile (>= n+2, i.e. n <= 22; multiple of 8)
constexpr int kBufRows = kRows + 2 * kMargin; // rows per activation buffer
constexpr int kChunks = kC / 8;               // 16-byte k-chunks per row
constexpr int kActBytes = kChunks * kBufRows * 16;
constexpr int kTaps = 9;
constexpr int kLayerWeightBytes = kTaps * kChunks * kC * 16;   // [tap][k-chunk][c_out][8 c_in] bf16
constexpr int kGroupBlocks = 4;               // M-blocks whose MMAs are interleaved (independent accumulators)
constexpr int kGroups = kBlocksM / kGroupBlocks;
constexpr int kMmaWarp = 8;                   // warps 0-7 epilogue (lane quarter = warp&3, channel half = warp>>2)
constexpr int kThreads = 32 * (kMmaWarp + 1); // warp 8: MMA issue + TMA
constexpr int kTmemCols = 32 * kBlocksM;      // 256
using namespace mnk_umma;

struct Smem {
    alignas(128) unsigned char act[2][kActBytes];
    alignas(128) unsigned char wts[2][kLayerWeightBytes];
    alignas(16) float head_w[3][kC];
    float head_b[4];
    alignas(8) unsigned long long mma_bar[kBlocksM];
    alignas(8) unsigned long long wts_bar[2];
    float head_part[128][3];                  // last layer: partial head dot products of the upper channel half
    unsigned int tmem_base;
};

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = 32
constexpr u32 kIdesc = umma_idesc_bf16(kC);

struct Params {
    int m, n, words, layers;          // layers = 1 + 2*blocks
    long long num_envs;
    int spc;                          // envs per CTA
    int pw, rs;                       // pixel-row stride n+1, env stride m*pw + pw + 1
    const u64* bits;                  // u64[2][words][num_envs]
    const u8* swap;                   // u8[num_envs] or NULL: 1 => the mover/agent is white, planes exchanged
    const unsigned char* weights;     // bf16 [layers][tap][k-chunk][c_out][8]
    const float* bias;                // f32 [layers][32]   (BN folded)
    const float* head_w;              // f32 [3][32]: policy ch0, policy ch1, value 1x1 conv
    const float* head_b;              // f32 [3]
    float* policy_feat;               // f32 [num_envs][2*cells]
    float* value_feat;                // f32 [num_envs][cells]
    int* error;                       // set to 1 on an mbarrier timeout
};

#ifndef MNK_EPI_PIPE
#define MNK_EPI_PIPE 1
#endif
#ifndef MNK_EPI_PACKSEL
#define MNK_EPI_PACKSEL 1
#endif

// bias (+ skip) + ReLU + bf16 store of NCH channels [ch0, ch0+NCH) of pixel row i; returns the fp32 values
template <int NCH>
MNK_DEV void epilogue_row(Smem& sm, const u32* acc, const float (&bias)[NCH], int out_buf, int i, int ch0, bool skip, bool valid,
                          bool store, float (&v)[NCH]) {
    constexpr int NKC = NCH / 8;
    uint4* out_row[NKC];
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc)
        out_row[kc] = reinterpret_cast<uint4*>(&sm.act[out_buf][0]) + (size_t)(ch0 / 8 + kc) * kBufRows + (kMargin + i);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) v[ch] = __uint_as_float(acc[ch]) + bias[ch];
    if (skip) {
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
            const uint4 rsd = *out_row[kc];
            const u32 w[4] = {rsd.x, rsd.y, rsd.z, rsd.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const float2 sk = act_unpack2(w[h]);
                v[kc * 8 + 2 * h] += sk.x;
                v[kc * 8 + 2 * h + 1] += sk.y;
            }
        }
    }
#if MNK_EPI_PACKSEL
    // guard rows are zeroed on the packed words (8 selects instead of 16); v[] of a guard row is not used by callers
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) v[ch] = fmaxf(v[ch], 0.0f);
    const u32 keep = valid ? 0xFFFFFFFFu : 0u;
#else
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) v[ch] = valid ? fmaxf(v[ch], 0.0f) : 0.0f;
    const u32 keep = 0xFFFFFFFFu;
#endif
    if (store) {
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
            u32 w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                w[h] = act_pack2(v[kc * 8 + 2 * h], v[kc * 8 + 2 * h + 1]) & keep;
            }
            *out_row[kc] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 2) resnet_tower_kernel(Params p) {
    // no pointer arithmetic through integers here: it would make every access a generic LD/ST instead of LDS/STS
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long env0 = (long long)blockIdx.x * p.spc;
    const int cells = p.m * p.n;
    const int envs_here = (int)min((long long)p.spc, p.num_envs - env0);

    // ---- one-time setup -----------------------------------------------------------------------------
    if (tid == 0) {
        for (int g = 0; g < kGroups; ++g) mbar_init(&sm.mma_bar[g], 1);
        mbar_init(&sm.wts_bar[0], 1);
        mbar_init(&sm.wts_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // first layer's weights: in flight while the tile is zeroed and the boards are decoded
        mbar_expect_tx(&sm.wts_bar[0], kLayerWeightBytes);
        tma_bulk_g2s(&sm.wts[0][0], p.weights, kLayerWeightBytes, &sm.wts_bar[0]);
    }
    if (warp == kMmaWarp) {   // TMEM allocation is a warp-wide operation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // Zero what no epilogue will write before it is read: k-chunks 0-1 of buffer 0 (the input layer reads
        // them; the decode below sets the stones) and the margins of every other k-chunk -- each epilogue
        // rewrites all kRows rows of its output buffer, zeros included.  Then the head weights.
        const uint4 zero = make_uint4(0, 0, 0, 0);
        uint4* a0 = reinterpret_cast<uint4*>(&sm.act[0][0]);
        for (int i = tid; i < 2 * kBufRows; i += kThreads) a0[i] = zero;
        for (int i = tid; i < 6 * 2 * kMargin; i += kThreads) {      // 6 remaining (buffer, k-chunk) planes x 2 margins
            const int plane = 2 + i / (2 * kMargin), r = i % (2 * kMargin);
            a0[plane * kBufRows + (r < kMargin ? r : kRows + r)] = zero;
        }
        for (int i = tid; i < 3 * kC; i += kThreads) (&sm.head_w[0][0])[i] = p.head_w[i];
        if (tid < 3) sm.head_b[tid] = p.head_b[tid];
    }
    __syncthreads();
    // input: the two canonical planes of each env into channels 0,1 of buffer A0 (k-chunk 0)
    for (int idx = tid; idx < envs_here * cells; idx += kThreads) {
        const int s = idx / cells, cell = idx - s * cells;
        const int r = cell / p.n, bit = cell + r;   // bit index in the guard-strided bitboard == row offset in the tile
        const long long e = env0 + s;
        const u64 wb = p.bits[(size_t)(bit >> 6) * p.num_envs + e];
        const u64 ww = p.bits[(size_t)(p.words + (bit >> 6)) * p.num_envs + e];
        const bool sw = p.swap != nullptr && p.swap[e] != 0;
        const u32 black = (u32)(wb >> (bit & 63)) & 1u, white = (u32)(ww >> (bit & 63)) & 1u;
        const u32 me = sw ? white : black, enemy = sw ? black : white;
        const int row = kMargin + s * p.rs + bit;
        reinterpret_cast<uint4*>(&sm.act[0][0])[row] = make_uint4(me * kActOne | (enemy * kActOne) << 16, 0, 0, 0);
    }
    // which of this thread's 8 pixel rows (one per M-block) are real board cells: fixed for all layers
    const int quarter = warp & 3, half = (warp >> 2) & 1;
    u32 valid_bits = 0;
    if (warp < kMmaWarp) {
        for (int j = 0; j < kBlocksM; ++j) {
            const int i = 128 * j + quarter * 32 + lane;
            const int s = i / p.rs, q = i - s * p.rs;
            const int r = q / p.pw, c = q - r * p.pw;
            if (s < envs_here && r < p.m && c < p.n) valid_bits |= 1u << j;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = sm.tmem_base;
    bool ok = true;
#ifdef MNK_TIMELINE   // debug build only (tools/timeline_tower.py): per-layer cycle stamps of one mid-grid CTA into error[1..]
    const long long t_origin = clock64();
    const bool stamp = p.error != nullptr && blockIdx.x == gridDim.x / 2 && lane == 0;
#define MNK_STAMP(slot) do { if (stamp) p.error[1 + 8 * L + (slot)] = (int)(clock64() - t_origin); } while (0)
#else
#define MNK_STAMP(slot) do { } while (0)
#endif

    int tap_rows[kTaps];   // row shift of each 3x3 tap in the guard-strided tile
#pragma unroll
    for (int tap = 0; tap < kTaps; ++tap) {
#ifdef MNK_EXP_TAP_OFF   // timing experiments only (wrong results)
        tap_rows[tap] = MNK_EXP_TAP_OFF;
#else
        tap_rows[tap] = (tap / 3 - 1) * p.pw + (tap % 3 - 1);
#endif
    }

    for (int L = 0; L < p.layers; ++L) {
        const int in_buf = (L & 1) ? 1 : 0, out_buf = in_buf ^ 1;
        const bool skip = (L >= 2) && ((L & 1) == 0);        // second conv of a residual block adds A1
        const bool last = (L == p.layers - 1);
        if (warp == kMmaWarp) {
            // The whole warp runs this (warp-uniform) control flow and ONE elected lane issues each
            // tcgen05 / TMA instruction: with a divergent `if (lane == 0)` region the compiler cannot
            // prove the descriptors uniform and wraps every MMA in an ELECT/R2UR waterfall loop
            // (~100 cycles per MMA on the issuing thread; ncu source page, profiles/).
            // next layer's weights into the other ring slot, ahead of this layer's MMAs (issuing the copy from inside
            // the MMA loop instead was 6 % slower: timeline in profiles/README.md)
            if (L + 1 < p.layers && elect_one()) {
                mbar_expect_tx(&sm.wts_bar[(L + 1) & 1], kLayerWeightBytes);
                tma_bulk_g2s(&sm.wts[(L + 1) & 1][0], p.weights + (size_t)(L + 1) * kLayerWeightBytes, kLayerWeightBytes,
                             &sm.wts_bar[(L + 1) & 1]);
            }
            ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.wts_bar[L & 1], (L >> 1) & 1)) != 0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            MNK_STAMP(0);   // MMA warp: issue loop starts
            // Descriptor words: the start-address field is (shared address >> 4); every tap / k-step / M-block
            // is that field plus a constant, so the fully unrolled loop below spends one add per operand.  (With
            // the index arithmetic in the loop the issuing thread needed ~60 cycles per MMA -- more than the 40
            // the tensor core takes to read the operands -- and the loop, not the pipe, set the pace: timeline
            // in profiles/README.md.)
            static_assert(kGroups == 1, "one interleaved group of M-blocks per layer");
            const u64 a_d0 = umma_desc(smem_u32(&sm.act[in_buf][0]) + kMargin * 16, kBufRows * 16, 128);   // row 0 of the tile, k-chunk 0
            const u64 b_d0 = umma_desc(smem_u32(&sm.wts[L & 1][0]), kC * 16, 128);
            const u32 a_lo0 = (u32)a_d0, a_hi = (u32)(a_d0 >> 32);      // low word: start-address field (bits 0-13) + LBO (bits 16-29)
            const u32 b_lo0 = (u32)b_d0, b_hi = (u32)(b_d0 >> 32);
            const bool two_ksteps = (L != 0);   // the input layer has 2 real channels: one K=16 step
#pragma unroll
            for (int tap = 0; tap < kTaps; ++tap) {
                const u32 a_tap = a_lo0 + (u32)tap_rows[tap];                 // rows are 16 B: row offset == field offset
                const u32 b_tap = b_lo0 + (u32)(tap * kChunks * kC);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        if (ks == 0 || two_ksteps) {
#pragma unroll
                            for (int jj = 0; jj < kBlocksM; ++jj)   // next M-block: +128 rows
                                umma_bf16_lohi(tmem_base + 32 * jj, a_tap + (u32)(2 * ks * kBufRows + 128 * jj), a_hi,
                                               b_tap + (u32)(2 * ks * kC), b_hi, kIdesc, (tap | ks) != 0);
                        }
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&sm.mma_bar[0]);
            __syncwarp();
            MNK_STAMP(1);   // MMA warp: all MMAs of the layer issued and committed
        } else {
            // this warp's 16 channels of the layer's folded bias, straight from global memory (L1-resident, issued
            // before the wait on the MMAs): as shared-memory loads they cost 2 % of the pipe the tensor core reads through
            float bias[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + L * kC + 16 * half) + q4);
                bias[4 * q4] = b4.x; bias[4 * q4 + 1] = b4.y; bias[4 * q4 + 2] = b4.z; bias[4 * q4 + 3] = b4.w;
            }
#if MNK_EPI_PIPE
            if (!last) {   // kGroups == 1: one wait, then the TMEM load of block j+1 is in flight under the arithmetic of block j
                static_assert(kGroups == 1, "pipelined epilogue assumes one commit per layer");
                ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.mma_bar[0], L & 1)) != 0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) MNK_STAMP(2);   // epilogue warp 0: MMAs complete (woken)
                const u32 t0 = tmem_base + ((u32)(quarter * 32) << 16) + 16 * half;
                u32 acc[2][16];
                tmem_ld16_issue(t0, acc[0]);
#pragma unroll
                for (int j = 0; j < kBlocksM; ++j) {
                    tmem_ld_wait(acc[j & 1]);
                    if (j + 1 < kBlocksM) tmem_ld16_issue(t0 + 32 * (j + 1), acc[(j + 1) & 1]);
                    float v[16];
                    epilogue_row<16>(sm, acc[j & 1], bias, out_buf, 128 * j + quarter * 32 + lane, 16 * half, skip,
                                     (valid_bits >> j) & 1u, true, v);
                }
                if (warp == 0) MNK_STAMP(3);   // epilogue warp 0: its four blocks done
            } else
#endif
            for (int j = 0; j < kBlocksM; ++j) {
                if (j % kGroupBlocks == 0) {
                    ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.mma_bar[j / kGroupBlocks], L & 1)) != 0;   // warp-uniform
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const int i = 128 * j + quarter * 32 + lane;        // pixel row of this thread
                const bool valid = (valid_bits >> j) & 1u;
                if (!last) {   // 8 warps: each takes 16 of the 32 channels of its TMEM lane quarter
                    u32 acc[16];
                    float v[16];
                    tmem_ld16(tmem_base + ((u32)(quarter * 32) << 16) + 32 * j + 16 * half, acc);
                    epilogue_row<16>(sm, acc, bias, out_buf, i, 16 * half, skip, valid, true, v);
                } else {   // last layer: no store; the 1x1 convolutions that open the two heads, from fp32 registers
                    u32 acc[16];
                    float v[16];
                    tmem_ld16(tmem_base + ((u32)(quarter * 32) << 16) + 32 * j + 16 * half, acc);
                    epilogue_row<16>(sm, acc, bias, out_buf, i, 16 * half, skip, valid, false, v);
                    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 16; ++ch) {
                        h0 = fmaf(v[ch], sm.head_w[0][16 * half + ch], h0);
                        h1 = fmaf(v[ch], sm.head_w[1][16 * half + ch], h1);
                        h2 = fmaf(v[ch], sm.head_w[2][16 * half + ch], h2);
                    }
                    float* part = sm.head_part[quarter * 32 + lane];
                    if (half == 1) { part[0] = h0; part[1] = h1; part[2] = h2; }
                    asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 epilogue warps
                    if (half == 0 && valid) {
                        h0 += part[0] + sm.head_b[0];
                        h1 += part[1] + sm.head_b[1];
                        h2 += part[2] + sm.head_b[2];
                        const int s = i / p.rs, q = i - s * p.rs;
                        const int r = q / p.pw, c = q - r * p.pw;
                        const long long e = env0 + s;
                        const int cell = r * p.n + c;
                        p.policy_feat[(size_t)e * 2 * cells + cell] = h0;
                        p.policy_feat[(size_t)e * 2 * cells + cells + cell] = h1;
                        p.value_feat[(size_t)e * cells + cell] = h2;
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
        }
        // layer boundary: epilogue stores -> visible to the async proxy (UMMA reads), TMEM reads done
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == kMmaWarp) MNK_STAMP(4);   // layer barrier passed
    }
    if (!ok && p.error != nullptr) atomicExch(p.error, 1);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}
}  // namespace rn

extern "C" int mnk_resnet_operand_dtype(void) { return mnk_umma::kActF16 ? 0 : 1; }

extern "C" int mnk_resnet_tower(const mnk_state_t* st, const uint8_t* swap, const void* weights, const float* bias,
                                const float* head_w, const float* head_b, int32_t blocks, float* policy_feat,
                                float* value_feat, int32_t* error, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (!weights || !bias || !head_w || !head_b || !policy_feat || !value_feat) return MNK_ERR_NULL;
    if (blocks < 1 || blocks > 8) return MNK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(bias)) & 15u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    rn::Params p;
    p.m = st->m; p.n = st->n; p.words = st->words; p.layers = 1 + 2 * blocks;
    p.num_envs = st->num_envs;
    p.pw = st->n + 1;
    p.rs = st->m * p.pw + p.pw + 1;
    p.spc = rn::kRows / p.rs;
    if (p.spc < 1 || p.pw + 1 > rn::kMargin) return MNK_ERR_GEOM;
    p.bits = reinterpret_cast<const u64*>(st->bits);
    p.swap = swap; p.weights = static_cast<const unsigned char*>(weights); p.bias = bias;
    p.head_w = head_w; p.head_b = head_b; p.policy_feat = policy_feat; p.value_feat = value_feat; p.error = error;
#ifndef MNK_EXTRA_SMEM   // experiments only: pad the request to force one CTA per SM
#define MNK_EXTRA_SMEM 0
#endif
    const size_t smem = sizeof(rn::Smem) + 128 + MNK_EXTRA_SMEM;
    static std::atomic<size_t> granted[kMaxDevices];
    if (int rc = mnk_optin_smem(rn::resnet_tower_kernel, smem, granted)) return rc;
    const unsigned grid = (unsigned)((st->num_envs + p.spc - 1) / p.spc);
    rn::resnet_tower_kernel<<<grid, rn::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return mnk_launch_status();
}
