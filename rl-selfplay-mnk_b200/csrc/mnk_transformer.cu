// mnk_transformer.cu -- the body of the reference's transformer policy / value networks on tcgen05.
//
// Reference: src/alg/architectures/transformer.py:7-92 with the parameters of configs.py:7-25 --
//   "transformer_b_s": embed_dim 56, 2 layers, 4 heads (head_dim 14), feed-forward 224
//   "transformer_b_l": embed_dim 96, 5 layers, 8 heads (head_dim 12), feed-forward 384
// cell_embed (Conv2d(2, D, 1)) + pos_embed, then `layers` x nn.TransformerEncoderLayer(batch_first, norm_first = True,
// ReLU, dropout 0): x += out_proj(softmax(Q K^T / sqrt(d)) V) with Q, K, V = in_proj(LayerNorm1(x)); x += linear2(ReLU(
// linear1(LayerNorm2(x)))); then the 1x1 Conv1d's that open the policy (D -> 2) and value (D -> 1) heads.
//
// One CTA owns 128 token rows = floor(128 / T) whole boards (T = m * n tokens, T <= 128) and runs ALL layers:
//   * the residual stream x lives in TENSOR MEMORY as fp32 (lane = token, D columns).  The output projection and the
//     second feed-forward GEMM accumulate straight into those columns (accumulate = 1): the residual add costs nothing;
//   * every GEMM is tcgen05.mma M = 128 with 16-bit operands (the library's operand type, fp16 by default) in the
//     K-major no-swizzle layout [k-chunk of 8][row][16 B] and fp32 accumulation: in_proj (N = 3 * heads * 16, head_dim
//     zero-padded to 16), per head S = Q K^T (N = 128 keys, K = 16) and O = P V (N = 16, K = 128 keys), out_proj,
//     linear1, linear2.  Attention is computed for all 128 token rows of the CTA at once and the softmax masks the keys of
//     other boards (block-diagonal), so boards never mix;
//   * LayerNorm, bias, ReLU, the softmax and the operand re-layouts run on 16 warps (TMEM lane quarter x column quarter:
//     the phases are latency-bound, 8 warps took twice as long),
//     each thread owning one token row of its part of the columns; V is written transposed (keys along K) for O = P V;
//   * weights stream through a two-slot ring of 1-D TMA bulk copies in consumption order (at most 36 KB per tile:
//     in_proj / linear1 split along N, linear2 along K at D = 96), requested two tiles ahead.
// Shared memory: operand buffer A (LayerNorm output / softmax P / attention output, 32 KB), operand buffer B (Q, K, V^T,
// later the feed-forward hidden activations, 56 / 96 KB), the weight ring, parameters.  Tensor memory: x at column 0,
// GEMM accumulators from column 96.
#include "mnk_dispatch.cuh"
#include "mnk_umma.cuh"

namespace tf {
using namespace mnk_umma;
constexpr int kRows = 128;
constexpr int kParts = 4;                      // column parts: each token row is shared by kParts threads
constexpr int kWorkers = 4 * kParts;           // warps 0-15: TMEM lane quarter = warp & 3, column part = warp >> 2
constexpr int kMmaWarp = kWorkers;             // warp 16: MMA issue + weight TMA
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr int kTmemCols = 512;
constexpr int kXCol = 0;                       // residual stream
constexpr int kWorkCol = 96;                   // GEMM accumulators (in_proj / scores + attention output / linear1)
constexpr int kPairBarrier = 1;                // named barrier: the 8 worker warps (row-sum exchanges between column halves)
constexpr float kLnEps = 1e-5f;
constexpr int kSC = kRows / kParts;            // score columns per thread in the softmax

template <int D, int NH>
struct Cfg {
    static constexpr int DP = (D + 15) / 16 * 16;          // embed_dim padded to a K step
    static constexpr int DH = D / NH;                      // real head_dim (14 / 12), padded to 16
    static constexpr int QP = NH * 16;                     // padded width of Q (= K = V)
    static constexpr int F = 4 * D;                        // feed-forward width (a multiple of 16 for both configs)
    static constexpr int NSPLIT = (3 * QP > 256) ? 2 : 1;  // in_proj / linear1 tiles along N; linear2 tiles along K
    static constexpr int NQKV = 3 * QP / NSPLIT;           // N of one in_proj MMA (192)
    static constexpr int NF1 = F / NSPLIT;                 // N of one linear1 MMA (224 / 192)
    static constexpr int KF2 = F / NSPLIT;                 // K of one linear2 tile
    static constexpr int kTilesPerLayer = 2 + 2 * NSPLIT + (NSPLIT - 1);   // in_proj, out_proj, linear1, linear2
    static constexpr int kTileQkv = DP * NQKV * 2, kTileO = QP * DP * 2, kTileF1 = DP * NF1 * 2, kTileF2 = KF2 * DP * 2;
    static constexpr int kTileMax = (kTileQkv > kTileF1 ? kTileQkv : kTileF1) > kTileF2 ? (kTileQkv > kTileF1 ? kTileQkv : kTileF1) : kTileF2;
    static constexpr int kLayerWeightBytes = NSPLIT * kTileQkv + kTileO + NSPLIT * kTileF1 + NSPLIT * kTileF2;
    static constexpr int kBufA = 16 * 2048;                // P needs 128 key columns = 16 chunks
    static constexpr int kQK = QP / 8 * 2048;              // Q (and K) operand bytes
    // V^T of one head: [key chunk (16)][16 dims][16 B] with the chunks 272 B apart instead of 256 (LBO is a descriptor field):
    // the 32 lanes of a warp -- four key chunks -- then scatter their 2-byte elements over distinct banks
    static constexpr int kVtChunk = 272, kVtHead = 16 * kVtChunk;
    static constexpr int kBufB = (2 * kQK + NH * kVtHead > F / 8 * 2048) ? 2 * kQK + NH * kVtHead : F / 8 * 2048;
    // per-layer fp32 parameters: ln1 g, b [DP] | in_proj bias [3 QP] | out_proj bias [DP] | ln2 g, b [DP] | b1 [F] | b2 [DP]
    static constexpr int oLn1 = 0, oBqkv = 2 * DP, oBo = oBqkv + 3 * QP, oLn2 = oBo + DP, oB1 = oLn2 + 2 * DP, oB2 = oB1 + F;
    static constexpr int kLayerParams = oB2 + DP;
    static_assert(DP <= kWorkCol && kWorkCol + 3 * QP <= kTmemCols && kWorkCol + F <= kTmemCols, "tensor memory plan");
    static_assert(kWorkCol + 256 + QP <= kTmemCols, "two score buffers + attention output");
    static_assert(DH <= 16 && DP % 16 == 0 && F % (16 * NSPLIT) == 0 && (3 * QP) % (16 * NSPLIT) == 0, "shapes");
};

template <int D, int NH>
struct Smem {
    using K = Cfg<D, NH>;
    alignas(128) unsigned char bufA[K::kBufA];
    alignas(128) unsigned char bufB[K::kBufB];
    alignas(128) unsigned char wts[2][K::kTileMax];
    alignas(16) float prm[K::kLayerParams];
    float emb[3][K::DP];                     // cell_embed weight (channel 0, channel 1), bias
    float head_w[3][K::DP];
    float head_b[4];
    float part[kParts][kRows][3];            // partial sums of the column parts (slots 0, 1; the head outputs use it flat)
    alignas(8) unsigned long long full_bar[2];
    unsigned long long mma_bar;
    unsigned int tmem_base;
};

struct Params {
    int m, n, words, layers, tokens;  // tokens = m * n
    long long num_envs;
    int spc;                          // boards per CTA = 128 / tokens
    const u64* bits;
    const u8* swap;
    const unsigned char* weights;     // op16 tiles in consumption order, kLayerWeightBytes per layer
    const float* layer_params;        // f32 [layers][kLayerParams]
    const float* embed;               // f32 [3][DP]: cell_embed weight of channel 0 / channel 1, bias
    const float* pos;                 // f32 [tokens][DP]
    const float* head_w;              // f32 [3][DP]
    const float* head_b;              // f32 [3]
    float* policy_feat;
    float* value_feat;
    int* error;
};

MNK_DEV void tmem_ld8(u32 taddr, u32 (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
                 :
                 : "memory");
}
// N columns (a multiple of 8) of this thread's TMEM lane: all loads in flight, ONE wait.  The empty asm statements pin every
// use of v[] behind the wait (volatile asm statements keep their order; each one redefines its register).
template <int N>
MNK_DEV void tmem_ld_cols(u32 taddr, u32 (&v)[N]) {
    static_assert(N % 8 == 0, "whole 8-column loads");
#pragma unroll
    for (int i = 0; i < N / 8; ++i)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[8 * i]), "=r"(v[8 * i + 1]), "=r"(v[8 * i + 2]), "=r"(v[8 * i + 3]), "=r"(v[8 * i + 4]),
                       "=r"(v[8 * i + 5]), "=r"(v[8 * i + 6]), "=r"(v[8 * i + 7])
                     : "r"(taddr + (u32)(8 * i)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(v[i]));
}
// f(c0, v8) for every 8-column chunk of W columns starting at taddr, 32 columns per load batch
template <int W, class F>
MNK_DEV void for_cols(u32 taddr, F&& f) {
    constexpr int B32 = W / 32, R = W % 32;
    static_assert(R % 8 == 0, "32-column batches and a tail of whole 8-column chunks");
#pragma unroll 1
    for (int b = 0; b < B32; ++b) {
        u32 v[32];
        tmem_ld_cols<32>(taddr + (u32)(32 * b), v);
#pragma unroll
        for (int c = 0; c < 4; ++c) f(32 * b + 8 * c, &v[8 * c]);
    }
    if constexpr (R != 0) {
        u32 v[R];
        tmem_ld_cols<R>(taddr + (u32)(32 * B32), v);
#pragma unroll
        for (int c = 0; c < R / 8; ++c) f(32 * B32 + 8 * c, &v[8 * c]);
    }
}
MNK_DEV void tmem_st8(u32 taddr, const u32 (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
MNK_DEV void tmem_st8p(u32 taddr, const u32* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
MNK_DEV float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
MNK_DEV void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
MNK_DEV void pair_sync() { asm volatile("bar.sync %0, %1;" ::"r"(kPairBarrier), "r"(32 * kWorkers) : "memory"); }

// all threads: phase boundary between CUDA-core writes (shared memory operands, tcgen05.st) and the next MMAs / loads
MNK_DEV void phase_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// one GEMM: D[128 x N] (+)= A[128 x 16 ksteps] * B^T, A chunks of 2048 B from a_base, B tile [k-chunk][N][16 B]
MNK_DEV void issue_gemm(u32 tmem_d, u32 a_base, int ksteps, u32 b_base, int N, bool accumulate) {
    const u32 idesc = umma_idesc_bf16(N);
    for (int ks = 0; ks < ksteps; ++ks) {
        const u64 a_d = umma_desc(a_base + (u32)(ks * 2 * 2048), 2048, 128);
        const u64 b_d = umma_desc(b_base + (u32)(ks * 2 * N * 16), (u32)(N * 16), 128);
        umma_bf16(tmem_d, a_d, b_d, idesc, (accumulate || ks != 0) ? 1u : 0u);
    }
}

template <int D, int NH>
__global__ void __launch_bounds__(kThreads, 1) transformer_body_kernel(Params p) {
    using K = Cfg<D, NH>;
    using S = Smem<D, NH>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    S& sm = *reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = warp < kWorkers;
    const int quarter = warp & 3, part = (warp >> 2) & (kParts - 1);
    const int row = quarter * 32 + lane;                  // token row = TMEM lane (workers)
    const int T = p.tokens;
    const long long env0 = (long long)blockIdx.x * p.spc;
    const int boards_here = (int)min((long long)p.spc, p.num_envs - env0);
    const int my_board = row / T, my_token = row - my_board * T;
    const bool row_valid = worker && my_board < boards_here;
    const bool warp_live = worker && quarter * 32 < boards_here * T;       // at least one real token row in this warp
    // keys of the boards that have a row in this warp (warp-uniform): the softmax skips the 8-column chunks outside
    const int u_lo = (quarter * 32 / T) * T, u_hi = min((quarter * 32 + 31) / T + 1, boards_here) * T;
    const bool one_board = u_hi - u_lo == T;                               // all real rows of the warp belong to one board
    const int total_tiles = p.layers * K::kTilesPerLayer;

    if (tid == 0) {
        mbar_init(&sm.full_bar[0], 1);
        mbar_init(&sm.full_bar[1], 1);
        mbar_init(&sm.mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 3 * K::DP; i += kThreads) {
        (&sm.emb[0][0])[i] = p.embed[i];
        (&sm.head_w[0][0])[i] = p.head_w[i];
    }
    if (tid < 3) sm.head_b[tid] = p.head_b[tid];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = sm.tmem_base;
    const u32 t_row = tmem + ((u32)(quarter * 32) << 16);     // this thread's TMEM lane
    bool ok = true;
    u32 mma_phase = 0;
#ifdef TF_TIMELINE   // debug build only (tools/timeline_tf.py): cycle stamps of one mid-grid CTA into error[1..]
    const long long t_origin = clock64();
    int stamp_idx = 0;
    const bool stamper = p.error != nullptr && blockIdx.x == gridDim.x / 2 && tid == 0;
#define TF_STAMP() do { if (stamper && stamp_idx < 120) p.error[1 + stamp_idx] = (int)(clock64() - t_origin); ++stamp_idx; } while (0)
#else
#define TF_STAMP() do { } while (0)
#endif

    // weight tiles in consumption order; tile i sits in ring slot i & 1
    auto tile_bytes = [&](int j) -> u32 {       // j = index within the layer
        if (j < K::NSPLIT) return K::kTileQkv;
        if (j == K::NSPLIT) return K::kTileO;
        if (j < 1 + 2 * K::NSPLIT) return K::kTileF1;
        return K::kTileF2;
    };
    auto tile_offset = [&](int i) -> size_t {
        const int L = i / K::kTilesPerLayer, j = i - L * K::kTilesPerLayer;
        size_t off = (size_t)L * K::kLayerWeightBytes;
        for (int q = 0; q < j; ++q) off += tile_bytes(q);
        return off;
    };
    auto request_tile = [&](int i) {            // MMA warp, one elected lane
        if (i < total_tiles) {
            const u32 bytes = tile_bytes(i % K::kTilesPerLayer);
            mbar_expect_tx(&sm.full_bar[i & 1], bytes);
            tma_bulk_g2s(&sm.wts[i & 1][0], p.weights + tile_offset(i), bytes, &sm.full_bar[i & 1]);
        }
    };
    int tile = 0;                                // next tile to consume (uniform across the CTA)
    if (warp == kMmaWarp && elect_one()) {
        request_tile(0);
        request_tile(1);
    }

    // ---- embedding: x = cell_embed(obs) + pos_embed, straight from the bitboards, into tensor memory -----------------
    if (worker) {
        float me = 0.f, enemy = 0.f;
        if (row_valid) {
            const int r = my_token / p.n, c = my_token - r * p.n, bit = my_token + r;
            const long long e = env0 + my_board;
            const u64 wb = p.bits[(size_t)(bit >> 6) * p.num_envs + e];
            const u64 ww = p.bits[(size_t)(p.words + (bit >> 6)) * p.num_envs + e];
            const bool sw = p.swap != nullptr && p.swap[e] != 0;
            const float black = (float)((wb >> (bit & 63)) & 1ull), white = (float)((ww >> (bit & 63)) & 1ull);
            me = sw ? white : black;
            enemy = sw ? black : white;
            (void)c;
        }
#pragma unroll
        for (int c8 = 0; c8 < K::DP / (8 * kParts); ++c8) {
            const int c0 = part * (K::DP / kParts) + 8 * c8;
            u32 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float x = 0.f;
                if (row_valid) x = fmaf(me, sm.emb[0][c0 + j], fmaf(enemy, sm.emb[1][c0 + j], sm.emb[2][c0 + j])) + __ldg(p.pos + (size_t)my_token * K::DP + c0 + j);
                v[j] = __float_as_uint(x);
            }
            tmem_st8(t_row + (u32)(kXCol + c0), v);
        }
        tmem_st_wait();
    }

    unsigned char* const bufA = &sm.bufA[0];
    unsigned char* const qbuf = &sm.bufB[0];
    unsigned char* const kbuf = &sm.bufB[K::kQK];
    unsigned char* const vtbuf = &sm.bufB[2 * K::kQK];
    unsigned char* const hbuf = &sm.bufB[0];
    const float scale = rsqrtf((float)K::DH);

    // LayerNorm of this thread's row (its column part) -> bufA as the next GEMM's A operand
    auto layer_norm = [&](const float* g, const float* b) {
        float x[K::DP / kParts];
        float sum = 0.f, sq = 0.f;
        {
            u32 v[K::DP / kParts];
            tmem_ld_cols<K::DP / kParts>(t_row + (u32)(kXCol + part * (K::DP / kParts)), v);
            float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};     // four independent chains
#pragma unroll
            for (int j = 0; j < K::DP / kParts; ++j) {
                x[j] = __uint_as_float(v[j]);
                s4[j & 3] += x[j];
                q4[j & 3] = fmaf(x[j], x[j], q4[j & 3]);
            }
            sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        }
        sm.part[part][row][0] = sum;
        sm.part[part][row][1] = sq;
        pair_sync();
        float tot = 0.f, tsq = 0.f;
#pragma unroll
        for (int q = 0; q < kParts; ++q) {
            tot += sm.part[q][row][0];
            tsq += sm.part[q][row][1];
        }
        const float mean = tot * (1.0f / D);
        const float rstd = rsqrtf(fmaxf(tsq * (1.0f / D) - mean * mean, 0.f) + kLnEps);
#pragma unroll
        for (int c8 = 0; c8 < K::DP / (8 * kParts); ++c8) {
            const int c0 = part * (K::DP / kParts) + 8 * c8;
            u32 w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const float y0 = (x[8 * c8 + 2 * h] - mean) * rstd * g[c0 + 2 * h] + b[c0 + 2 * h];
                const float y1 = (x[8 * c8 + 2 * h + 1] - mean) * rstd * g[c0 + 2 * h + 1] + b[c0 + 2 * h + 1];
                w[h] = act_pack2(y0, y1);          // padded columns: g = b = 0 -> 0
            }
            *reinterpret_cast<uint4*>(bufA + ((size_t)(c0 / 8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    // x[row][this thread's columns] += bias (the bias of a GEMM that accumulates into the residual stream)
    auto add_bias_to_x = [&](const float* bias) {
        const int c0 = part * (K::DP / kParts);
        u32 v[K::DP / kParts];
        tmem_ld_cols<K::DP / kParts>(t_row + (u32)(kXCol + c0), v);
#pragma unroll
        for (int j = 0; j < K::DP / kParts; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + bias[c0 + j]);
#pragma unroll
        for (int c8 = 0; c8 < K::DP / (8 * kParts); ++c8) tmem_st8p(t_row + (u32)(kXCol + c0 + 8 * c8), &v[8 * c8]);
        tmem_st_wait();
    };
    // all threads: wait for the commit of the MMAs just issued
    auto wait_mma = [&]() {
        ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.mma_bar, mma_phase)) != 0;
        mma_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    // MMA warp: wait for weight tile `i`
    auto wait_tile = [&](int i) {
        ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.full_bar[i & 1], (u32)(i >> 1) & 1u)) != 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    for (int L = 0; L < p.layers; ++L) {
        // ---- this layer's parameters; LayerNorm1 -> bufA ---------------------------------------------------------------
        __syncthreads();                       // every reader of the previous layer's parameters is done
        for (int i = tid; i < K::kLayerParams; i += kThreads) sm.prm[i] = __ldg(p.layer_params + (size_t)L * K::kLayerParams + i);
        __syncthreads();
        if (worker) layer_norm(&sm.prm[K::oLn1], &sm.prm[K::oLn1 + K::DP]);
        phase_sync();
        TF_STAMP();
        // ---- in_proj: [Q | K | V] = LN1(x) Wqkv^T ------------------------------------------------------------------------
        if (warp == kMmaWarp) {
            for (int j = 0; j < K::NSPLIT; ++j) {
                wait_tile(tile + j);
                if (elect_one())
                    issue_gemm(tmem + (u32)(kWorkCol + j * K::NQKV), smem_u32(bufA), K::DP / 16, smem_u32(&sm.wts[(tile + j) & 1][0]), K::NQKV, false);
                __syncwarp();
            }
            if (elect_one()) umma_commit(&sm.mma_bar);
            __syncwarp();
        }
        wait_mma();
        TF_STAMP();
        if (warp == kMmaWarp && elect_one())
            for (int j = 0; j < K::NSPLIT; ++j) request_tile(tile + j + 2);
        tile += K::NSPLIT;
        // ---- + bias, Q scaled, operands for the attention: Q, K as [chunk][token][16 B], V transposed ------------------
        if (worker) {
            constexpr int W = 3 * K::QP / kParts;     // columns per part
            for_cols<W>(t_row + (u32)(kWorkCol + part * W), [&](int rel, const u32* v) {
                const int c0 = part * W + rel;
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(v[j]) + sm.prm[K::oBqkv + c0 + j];
                const int which = c0 / K::QP, within = c0 - which * K::QP;
                if (which < 2) {
                    const float s = which == 0 ? scale : 1.0f;
                    u32 w[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) w[h] = act_pack2(y[2 * h] * s, y[2 * h + 1] * s);
                    unsigned char* dst = (which == 0 ? qbuf : kbuf) + ((size_t)(within / 8) * kRows + row) * 16;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
                    const int head = within / 16, d0 = within & 15;
                    unsigned char* dst = vtbuf + (size_t)head * K::kVtHead + (size_t)(row >> 3) * K::kVtChunk + d0 * 16 + (row & 7) * 2;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<unsigned short*>(dst + j * 16) = (unsigned short)(act_pack2(y[j], 0.f) & 0xFFFFu);
                }
            });
        }
        phase_sync();
        TF_STAMP();
        // ---- attention, head by head: S = Q K^T -> softmax over this board's keys -> O = P V -------------------------------
#ifndef TF_EXP_HEADS          // timing experiments only (wrong results): how many heads run
#define TF_EXP_HEADS NH
#endif
        // The scores of head h + 1 are issued together with O_h = P_h V_h (two score buffers in tensor memory): one MMA
        // round trip per head instead of two.
        if (warp == kMmaWarp) {
            if (elect_one()) {
                issue_gemm(tmem + (u32)kWorkCol, smem_u32(qbuf), 1, smem_u32(kbuf), 128, false);
                umma_commit(&sm.mma_bar);
            }
            __syncwarp();
        }
        wait_mma();
        TF_STAMP();
#pragma unroll 1
        for (int h = 0; h < TF_EXP_HEADS; ++h) {
            const u32 s_col = (u32)(kWorkCol + 128 * (h & 1));          // this head's score columns
#ifdef TF_EXP_NO_SOFTMAX
            if (false) {
#else
            if (worker && !warp_live) {
#endif
                // every row of this warp is padding: P = 0 for them, no arithmetic (the exchanges still take place)
#ifndef TF_EXP_NO_PAIRSYNC
                pair_sync();
                pair_sync();
#endif
#pragma unroll
                for (int c8 = 0; c8 < kSC / 8; ++c8)
                    *reinterpret_cast<uint4*>(bufA + ((size_t)(kSC / 8 * part + c8) * kRows + row) * 16) = make_uint4(0, 0, 0, 0);
#ifdef TF_EXP_NO_SOFTMAX
            } else if (false) {
#else
            } else if (worker) {
#endif
                float e[kSC];
                // this board's keys; a padding row of a single-board warp keeps that board's keys (its result is never used,
                // it only has to stay finite)
                const int lo = one_board ? u_lo : my_board * T, hi = one_board ? u_hi : (row_valid ? lo + T : lo);
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // four independent chains
                {
                    u32 v[kSC];
                    tmem_ld_cols<kSC>(t_row + s_col + (u32)(kSC * part), v);
#pragma unroll
                    for (int c8 = 0; c8 < kSC / 8; ++c8) {
                        // keys of no board that has a row in this warp: nothing to compute (warp-uniform branch) -- at 9x9 a
                        // CTA holds one board, 47 of its 128 key columns and 47 of its rows are padding
                        const int col0 = kSC * part + 8 * c8;
                        if (one_board && col0 >= u_lo && col0 + 8 <= u_hi) {
                            // every row of the warp sees all 8 keys (its padding rows may see anything finite): no masks
#pragma unroll
                            for (int j = 8 * c8; j < 8 * c8 + 8; ++j) {
                                e[j] = __uint_as_float(v[j]);
                                m4[j & 3] = fmaxf(m4[j & 3], e[j]);
                            }
                        } else if (col0 + 8 > u_lo && col0 < u_hi) {
#pragma unroll
                            for (int j = 8 * c8; j < 8 * c8 + 8; ++j) {
                                const int col = kSC * part + j;
                                e[j] = (col >= lo && col < hi) ? __uint_as_float(v[j]) : -INFINITY;
                                m4[j & 3] = fmaxf(m4[j & 3], e[j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 8 * c8; j < 8 * c8 + 8; ++j) e[j] = -INFINITY;
                        }
                    }
                }
                float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                sm.part[part][row][0] = mx;
#ifndef TF_EXP_NO_PAIRSYNC
                pair_sync();
#endif
                mx = fmaxf(fmaxf(sm.part[0][row][0], sm.part[1][row][0]), fmaxf(sm.part[2][row][0], sm.part[3][row][0]));
                if (mx == -INFINITY) mx = 0.f;                                   // a padding row with no key at all: P = 0, not NaN
                const float mxl = mx * 1.4426950408889634f;                      // exp(s - mx) = exp2(s * log2(e) - mx * log2(e))
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c8 = 0; c8 < kSC / 8; ++c8) {
                    const int col0 = kSC * part + 8 * c8;
                    if (col0 + 8 > u_lo && col0 < u_hi) {
#pragma unroll
                        for (int j = 8 * c8; j < 8 * c8 + 8; ++j) {
#ifdef TF_EXP_NO_EXP
                            e[j] = (e[j] == -INFINITY) ? 0.f : fmaf(e[j], 1.4426950408889634f, -mxl);
#else
                            e[j] = fast_exp2(fmaf(e[j], 1.4426950408889634f, -mxl));      // ex2.approx(-inf) = +0: masked keys vanish
#endif
                            s4[j & 3] += e[j];
                        }
                    } else {
#pragma unroll
                        for (int j = 8 * c8; j < 8 * c8 + 8; ++j) e[j] = 0.f;
                    }
                }
                float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                sm.part[part][row][1] = sum;
#ifndef TF_EXP_NO_PAIRSYNC
                pair_sync();
#endif
                sum = (sm.part[0][row][1] + sm.part[1][row][1]) + (sm.part[2][row][1] + sm.part[3][row][1]);
                const float inv = sum > 0.f ? __frcp_rn(sum) : 0.f;
#pragma unroll
                for (int c8 = 0; c8 < kSC / 8; ++c8) {
                    u32 w[4] = {0u, 0u, 0u, 0u};
                    const int col0 = kSC * part + 8 * c8;
                    if (col0 + 8 > u_lo && col0 < u_hi) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) w[q] = act_pack2(e[8 * c8 + 2 * q] * inv, e[8 * c8 + 2 * q + 1] * inv);
                    }
                    *reinterpret_cast<uint4*>(bufA + ((size_t)(kSC / 8 * part + c8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            phase_sync();
            TF_STAMP();
            if (warp == kMmaWarp) {
                if (elect_one()) {
                    // O_h[128 x 16] = P[128 x 128 keys] V_h: B = V^T_h as [key chunk][16 dims][16 B]
                    const u32 idesc = umma_idesc_bf16(16);
#ifndef TF_EXP_PV_KSTEPS
#define TF_EXP_PV_KSTEPS 8
#endif
                    for (int ks = 0; ks < TF_EXP_PV_KSTEPS; ++ks) {
                        const u64 a_d = umma_desc(smem_u32(bufA) + (u32)(ks * 2 * 2048), 2048, 128);
                        const u64 b_d = umma_desc(smem_u32(vtbuf) + (u32)(h * K::kVtHead + ks * 2 * K::kVtChunk), K::kVtChunk, 128);
                        umma_bf16(tmem + (u32)(kWorkCol + 256 + 16 * h), a_d, b_d, idesc, ks != 0);
                    }
                    if (h + 1 < NH)
                        issue_gemm(tmem + (u32)(kWorkCol + 128 * ((h + 1) & 1)), smem_u32(qbuf) + (u32)((h + 1) * 2 * 2048), 1,
                                   smem_u32(kbuf) + (u32)((h + 1) * 2 * 2048), 128, false);
                    umma_commit(&sm.mma_bar);
                }
                __syncwarp();
            }
            wait_mma();            // P (bufA) and the score columns are free again
            TF_STAMP();
        }
        // ---- attention output -> bufA as the out_proj operand; x += out_proj bias ------------------------------------------
        if (worker) {
            for_cols<K::QP / kParts>(t_row + (u32)(kWorkCol + 256 + part * (K::QP / kParts)), [&](int rel, const u32* v) {
                const int c0 = part * (K::QP / kParts) + rel;
                u32 w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = act_pack2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
                *reinterpret_cast<uint4*>(bufA + ((size_t)(c0 / 8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            });
            add_bias_to_x(&sm.prm[K::oBo]);
        }
        phase_sync();
        TF_STAMP();
        // ---- out_proj, accumulated into the residual stream --------------------------------------------------------------------
        if (warp == kMmaWarp) {
            wait_tile(tile);
            if (elect_one()) {
                issue_gemm(tmem + (u32)kXCol, smem_u32(bufA), K::QP / 16, smem_u32(&sm.wts[tile & 1][0]), K::DP, true);
                umma_commit(&sm.mma_bar);
            }
            __syncwarp();
        }
        wait_mma();
        TF_STAMP();
        if (warp == kMmaWarp && elect_one()) request_tile(tile + 2);
        tile += 1;
        // ---- LayerNorm2 -> bufA; linear1 ------------------------------------------------------------------------------------------
        if (worker) layer_norm(&sm.prm[K::oLn2], &sm.prm[K::oLn2 + K::DP]);
        phase_sync();
        TF_STAMP();
        if (warp == kMmaWarp) {
            for (int j = 0; j < K::NSPLIT; ++j) {
                wait_tile(tile + j);
                if (elect_one())
                    issue_gemm(tmem + (u32)(kWorkCol + j * K::NF1), smem_u32(bufA), K::DP / 16, smem_u32(&sm.wts[(tile + j) & 1][0]), K::NF1, false);
                __syncwarp();
            }
            if (elect_one()) umma_commit(&sm.mma_bar);
            __syncwarp();
        }
        wait_mma();
        TF_STAMP();
        if (warp == kMmaWarp && elect_one())
            for (int j = 0; j < K::NSPLIT; ++j) request_tile(tile + j + 2);
        tile += K::NSPLIT;
        // ---- + bias, ReLU -> hidden activations (bufB) as the linear2 operand; x += linear2 bias -----------------------------
        if (worker) {
            constexpr int W = K::F / kParts;
            for_cols<W>(t_row + (u32)(kWorkCol + part * W), [&](int rel, const u32* v) {
                const int c0 = part * W + rel;
                u32 w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    w[q] = act_pack2(fmaxf(__uint_as_float(v[2 * q]) + sm.prm[K::oB1 + c0 + 2 * q], 0.f),
                                     fmaxf(__uint_as_float(v[2 * q + 1]) + sm.prm[K::oB1 + c0 + 2 * q + 1], 0.f));
                *reinterpret_cast<uint4*>(hbuf + ((size_t)(c0 / 8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            });
            add_bias_to_x(&sm.prm[K::oB2]);
        }
        phase_sync();
        TF_STAMP();
        // ---- linear2, accumulated into the residual stream ----------------------------------------------------------------------
        if (warp == kMmaWarp) {
            for (int j = 0; j < K::NSPLIT; ++j) {
                wait_tile(tile + j);
                if (elect_one())
                    issue_gemm(tmem + (u32)kXCol, smem_u32(hbuf) + (u32)(j * (K::KF2 / 8) * 2048), K::KF2 / 16, smem_u32(&sm.wts[(tile + j) & 1][0]), K::DP, true);
                __syncwarp();
            }
            if (elect_one()) umma_commit(&sm.mma_bar);
            __syncwarp();
        }
        wait_mma();
        TF_STAMP();
        if (warp == kMmaWarp && elect_one())
            for (int j = 0; j < K::NSPLIT; ++j) request_tile(tile + j + 2);
        tile += K::NSPLIT;
    }

    // ---- the 1x1 convolutions that open the two heads (Conv1d(D, 2, 1), Conv1d(D, 1, 1)) -------------------------------------
    if (worker) {
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        {
            const int c0 = part * (K::DP / kParts);
            u32 v[K::DP / kParts];
            tmem_ld_cols<K::DP / kParts>(t_row + (u32)(kXCol + c0), v);
#pragma unroll
            for (int j = 0; j < K::DP / kParts; ++j) {
                const float x = __uint_as_float(v[j]);
                h0 = fmaf(x, sm.head_w[0][c0 + j], h0);
                h1 = fmaf(x, sm.head_w[1][c0 + j], h1);
                h2 = fmaf(x, sm.head_w[2][c0 + j], h2);
            }
        }
        pair_sync();                            // the softmax / LayerNorm readers of `part` are done
        float* flat = &sm.part[0][0][0];        // `part` as a flat array: [3 outputs][kParts - 1 parts][128 rows] = 1,152 of its 1,536 floats
        if (part != 0) {
            flat[(0 * (kParts - 1) + part - 1) * kRows + row] = h0;
            flat[(1 * (kParts - 1) + part - 1) * kRows + row] = h1;
            flat[(2 * (kParts - 1) + part - 1) * kRows + row] = h2;
        }
        pair_sync();
        if (part == 0 && row_valid) {
#pragma unroll
            for (int q = 0; q < kParts - 1; ++q) {
                h0 += flat[(0 * (kParts - 1) + q) * kRows + row];
                h1 += flat[(1 * (kParts - 1) + q) * kRows + row];
                h2 += flat[(2 * (kParts - 1) + q) * kRows + row];
            }
            const long long e = env0 + my_board;
            p.policy_feat[(size_t)e * 2 * T + my_token] = h0 + sm.head_b[0];
            p.policy_feat[(size_t)e * 2 * T + T + my_token] = h1 + sm.head_b[1];
            p.value_feat[(size_t)e * T + my_token] = h2 + sm.head_b[2];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && p.error != nullptr) atomicMax(p.error, 8);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

template <int D, int NH>
static int launch(Params& p, cudaStream_t s) {
    const size_t smem = sizeof(Smem<D, NH>) + 128;
    if (smem > 227 * 1024) return MNK_ERR_GEOM;
    static std::atomic<size_t> granted[kMaxDevices];
    if (int rc = mnk_optin_smem(transformer_body_kernel<D, NH>, smem, granted)) return rc;
    const unsigned grid = (unsigned)((p.num_envs + p.spc - 1) / p.spc);
    transformer_body_kernel<D, NH><<<grid, kThreads, smem, s>>>(p);
    return mnk_launch_status();
}
}  // namespace tf

extern "C" int64_t mnk_transformer_layer_weight_bytes(int32_t embed_dim, int32_t heads) {
    if (embed_dim == 56 && heads == 4) return tf::Cfg<56, 4>::kLayerWeightBytes;
    if (embed_dim == 96 && heads == 8) return tf::Cfg<96, 8>::kLayerWeightBytes;
    return MNK_ERR_ARG;
}

extern "C" int64_t mnk_transformer_layer_params(int32_t embed_dim, int32_t heads) {
    if (embed_dim == 56 && heads == 4) return tf::Cfg<56, 4>::kLayerParams;
    if (embed_dim == 96 && heads == 8) return tf::Cfg<96, 8>::kLayerParams;
    return MNK_ERR_ARG;
}

extern "C" int mnk_transformer_body(const mnk_state_t* st, const uint8_t* swap, int32_t embed_dim, int32_t heads, int32_t layers,
                                    const void* weights, const float* layer_params, const float* embed, const float* pos,
                                    const float* head_w, const float* head_b, float* policy_feat, float* value_feat,
                                    int32_t* error, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (!weights || !layer_params || !embed || !pos || !head_w || !head_b || !policy_feat || !value_feat) return MNK_ERR_NULL;
    if (layers < 1 || layers > 16) return MNK_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(weights) & 15u) return MNK_ERR_ALIGN;
    const int tokens = st->m * st->n;
    if (tokens > tf::kRows) return MNK_ERR_GEOM;
    if (st->num_envs == 0) return MNK_OK;
    tf::Params p;
    p.m = st->m; p.n = st->n; p.words = st->words; p.layers = layers; p.tokens = tokens;
    p.num_envs = st->num_envs; p.spc = tf::kRows / tokens;
    p.bits = reinterpret_cast<const u64*>(st->bits); p.swap = swap;
    p.weights = static_cast<const unsigned char*>(weights); p.layer_params = layer_params; p.embed = embed; p.pos = pos;
    p.head_w = head_w; p.head_b = head_b; p.policy_feat = policy_feat; p.value_feat = value_feat; p.error = error;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (embed_dim == 56 && heads == 4) return tf::launch<56, 4>(p, s);
    if (embed_dim == 96 && heads == 8) return tf::launch<96, 8>(p, s);
    return MNK_ERR_ARG;
}
