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
//   * LayerNorm, bias, ReLU, the softmax and the operand re-layouts run on 8 warps (TMEM lane quarter x column half),
//     each thread owning one token row of its half of the columns; V is written transposed (keys along K) for O = P V;
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
constexpr int kWorkers = 8;                    // warps 0-7: TMEM lane quarter = warp & 3, column half = warp >> 2
constexpr int kMmaWarp = kWorkers;             // warp 8: MMA issue + weight TMA
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr int kTmemCols = 512;
constexpr int kXCol = 0;                       // residual stream
constexpr int kWorkCol = 96;                   // GEMM accumulators (in_proj / scores + attention output / linear1)
constexpr int kPairBarrier = 1;                // named barrier: the 8 worker warps (row-sum exchanges between column halves)
constexpr float kLnEps = 1e-5f;

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
    static constexpr int kBufB = (2 * kQK + NH * 4096 > F / 8 * 2048) ? 2 * kQK + NH * 4096 : F / 8 * 2048;
    // per-layer fp32 parameters: ln1 g, b [DP] | in_proj bias [3 QP] | out_proj bias [DP] | ln2 g, b [DP] | b1 [F] | b2 [DP]
    static constexpr int oLn1 = 0, oBqkv = 2 * DP, oBo = oBqkv + 3 * QP, oLn2 = oBo + DP, oB1 = oLn2 + 2 * DP, oB2 = oB1 + F;
    static constexpr int kLayerParams = oB2 + DP;
    static_assert(DP <= kWorkCol && kWorkCol + 3 * QP <= kTmemCols && kWorkCol + F <= kTmemCols, "tensor memory plan");
    static_assert(kWorkCol + 128 + QP <= kTmemCols, "scores + attention output");
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
    float part[2][kRows][2];                 // column-half partial sums (two slots)
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
MNK_DEV void tmem_st8(u32 taddr, const u32 (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
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
    const int quarter = warp & 3, half = (warp >> 2) & 1;
    const int row = quarter * 32 + lane;                  // token row = TMEM lane (workers)
    const int T = p.tokens;
    const long long env0 = (long long)blockIdx.x * p.spc;
    const int boards_here = (int)min((long long)p.spc, p.num_envs - env0);
    const int my_board = row / T, my_token = row - my_board * T;
    const bool row_valid = worker && my_board < boards_here;
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
        for (int c8 = 0; c8 < K::DP / 16; ++c8) {
            const int c0 = half * (K::DP / 2) + 8 * c8;
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

    // LayerNorm of this thread's row (its column half) -> bufA as the next GEMM's A operand
    auto layer_norm = [&](const float* g, const float* b) {
        float x[K::DP / 2];
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < K::DP / 16; ++c8) {
            u32 v[8];
            tmem_ld8(t_row + (u32)(kXCol + half * (K::DP / 2) + 8 * c8), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x[8 * c8 + j] = __uint_as_float(v[j]);
                sum += x[8 * c8 + j];
                sq = fmaf(x[8 * c8 + j], x[8 * c8 + j], sq);
            }
        }
        sm.part[half][row][0] = sum;
        sm.part[half][row][1] = sq;
        pair_sync();
        const float tot = sm.part[0][row][0] + sm.part[1][row][0], tsq = sm.part[0][row][1] + sm.part[1][row][1];
        const float mean = tot * (1.0f / D);
        const float rstd = rsqrtf(fmaxf(tsq * (1.0f / D) - mean * mean, 0.f) + kLnEps);
#pragma unroll
        for (int c8 = 0; c8 < K::DP / 16; ++c8) {
            const int c0 = half * (K::DP / 2) + 8 * c8;
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
    // x[row][half columns] += bias (the bias of a GEMM that accumulates into the residual stream)
    auto add_bias_to_x = [&](const float* bias) {
#pragma unroll
        for (int c8 = 0; c8 < K::DP / 16; ++c8) {
            const int c0 = half * (K::DP / 2) + 8 * c8;
            u32 v[8];
            tmem_ld8(t_row + (u32)(kXCol + c0), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + bias[c0 + j]);
            tmem_st8(t_row + (u32)(kXCol + c0), v);
        }
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
        if (warp == kMmaWarp && elect_one())
            for (int j = 0; j < K::NSPLIT; ++j) request_tile(tile + j + 2);
        tile += K::NSPLIT;
        // ---- + bias, Q scaled, operands for the attention: Q, K as [chunk][token][16 B], V transposed ------------------
        if (worker) {
            constexpr int W = 3 * K::QP / 2;     // columns per half
#pragma unroll 1
            for (int c8 = 0; c8 < W / 8; ++c8) {
                const int c0 = half * W + 8 * c8;
                u32 v[8];
                tmem_ld8(t_row + (u32)(kWorkCol + c0), v);
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
                    unsigned char* dst = vtbuf + (size_t)head * 4096 + ((size_t)(row >> 3) * 16 + d0) * 16 + (row & 7) * 2;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<unsigned short*>(dst + j * 16) = (unsigned short)(act_pack2(y[j], 0.f) & 0xFFFFu);
                }
            }
        }
        phase_sync();
        // ---- attention, head by head: S = Q K^T -> softmax over this board's keys -> O = P V -------------------------------
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            if (warp == kMmaWarp) {
                if (elect_one()) {
                    issue_gemm(tmem + (u32)kWorkCol, smem_u32(qbuf) + (u32)(h * 2 * 2048), 1, smem_u32(kbuf) + (u32)(h * 2 * 2048), 128, false);
                    umma_commit(&sm.mma_bar);
                }
                __syncwarp();
            }
            wait_mma();
            if (worker) {
                float e[64];
                float mx = -INFINITY;
                const int lo = my_board * T, hi = row_valid ? lo + T : lo;       // this board's keys
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    u32 v[8];
                    tmem_ld8(t_row + (u32)(kWorkCol + 64 * half + 8 * c8), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = 64 * half + 8 * c8 + j;
                        e[8 * c8 + j] = (col >= lo && col < hi) ? __uint_as_float(v[j]) : -INFINITY;
                        mx = fmaxf(mx, e[8 * c8 + j]);
                    }
                }
                sm.part[half][row][0] = mx;
                pair_sync();
                mx = fmaxf(sm.part[0][row][0], sm.part[1][row][0]);
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    e[j] = (e[j] == -INFINITY) ? 0.f : __expf(e[j] - mx);
                    sum += e[j];
                }
                sm.part[half][row][1] = sum;
                pair_sync();
                sum = sm.part[0][row][1] + sm.part[1][row][1];
                const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    u32 w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) w[q] = act_pack2(e[8 * c8 + 2 * q] * inv, e[8 * c8 + 2 * q + 1] * inv);
                    *reinterpret_cast<uint4*>(bufA + ((size_t)(8 * half + c8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            phase_sync();
            if (warp == kMmaWarp) {
                if (elect_one()) {
                    // O_h[128 x 16] = P[128 x 128 keys] V_h: B = V^T_h as [key chunk][16 dims][16 B]
                    const u32 idesc = umma_idesc_bf16(16);
                    for (int ks = 0; ks < 8; ++ks) {
                        const u64 a_d = umma_desc(smem_u32(bufA) + (u32)(ks * 2 * 2048), 2048, 128);
                        const u64 b_d = umma_desc(smem_u32(vtbuf) + (u32)(h * 4096 + ks * 2 * 256), 256, 128);
                        umma_bf16(tmem + (u32)(kWorkCol + 128 + 16 * h), a_d, b_d, idesc, ks != 0);
                    }
                    umma_commit(&sm.mma_bar);
                }
                __syncwarp();
            }
            wait_mma();            // P (bufA) and the score columns are free again
        }
        // ---- attention output -> bufA as the out_proj operand; x += out_proj bias ------------------------------------------
        if (worker) {
#pragma unroll 1
            for (int c8 = 0; c8 < K::QP / 16; ++c8) {
                const int c0 = half * (K::QP / 2) + 8 * c8;
                u32 v[8];
                tmem_ld8(t_row + (u32)(kWorkCol + 128 + c0), v);
                u32 w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = act_pack2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
                *reinterpret_cast<uint4*>(bufA + ((size_t)(c0 / 8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            add_bias_to_x(&sm.prm[K::oBo]);
        }
        phase_sync();
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
        if (warp == kMmaWarp && elect_one()) request_tile(tile + 2);
        tile += 1;
        // ---- LayerNorm2 -> bufA; linear1 ------------------------------------------------------------------------------------------
        if (worker) layer_norm(&sm.prm[K::oLn2], &sm.prm[K::oLn2 + K::DP]);
        phase_sync();
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
        if (warp == kMmaWarp && elect_one())
            for (int j = 0; j < K::NSPLIT; ++j) request_tile(tile + j + 2);
        tile += K::NSPLIT;
        // ---- + bias, ReLU -> hidden activations (bufB) as the linear2 operand; x += linear2 bias -----------------------------
        if (worker) {
            constexpr int W = K::F / 2;
#pragma unroll 1
            for (int c8 = 0; c8 < W / 8; ++c8) {
                const int c0 = half * W + 8 * c8;
                u32 v[8];
                tmem_ld8(t_row + (u32)(kWorkCol + c0), v);
                u32 w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    w[q] = act_pack2(fmaxf(__uint_as_float(v[2 * q]) + sm.prm[K::oB1 + c0 + 2 * q], 0.f),
                                     fmaxf(__uint_as_float(v[2 * q + 1]) + sm.prm[K::oB1 + c0 + 2 * q + 1], 0.f));
                *reinterpret_cast<uint4*>(hbuf + ((size_t)(c0 / 8) * kRows + row) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            add_bias_to_x(&sm.prm[K::oB2]);
        }
        phase_sync();
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
        if (warp == kMmaWarp && elect_one())
            for (int j = 0; j < K::NSPLIT; ++j) request_tile(tile + j + 2);
        tile += K::NSPLIT;
    }

    // ---- the 1x1 convolutions that open the two heads (Conv1d(D, 2, 1), Conv1d(D, 1, 1)) -------------------------------------
    if (worker) {
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < K::DP / 16; ++c8) {
            const int c0 = half * (K::DP / 2) + 8 * c8;
            u32 v[8];
            tmem_ld8(t_row + (u32)(kXCol + c0), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = __uint_as_float(v[j]);
                h0 = fmaf(x, sm.head_w[0][c0 + j], h0);
                h1 = fmaf(x, sm.head_w[1][c0 + j], h1);
                h2 = fmaf(x, sm.head_w[2][c0 + j], h2);
            }
        }
        pair_sync();                            // the softmax / LayerNorm readers of `part` are done
        float* mine = &sm.part[0][0][0] + 3 * row;     // `part` as a flat array: three floats per row
        if (half == 1) { mine[0] = h0; mine[1] = h1; mine[2] = h2; }
        pair_sync();
        if (half == 0 && row_valid) {
            const long long e = env0 + my_board;
            p.policy_feat[(size_t)e * 2 * T + my_token] = h0 + mine[0] + sm.head_b[0];
            p.policy_feat[(size_t)e * 2 * T + T + my_token] = h1 + mine[1] + sm.head_b[1];
            p.value_feat[(size_t)e * T + my_token] = h2 + mine[2] + sm.head_b[2];
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
