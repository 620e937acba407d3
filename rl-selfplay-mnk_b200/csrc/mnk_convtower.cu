// mnk_convtower.cu -- the convolutional body of the reference's WIDER networks on tcgen05.
//
// Reference architectures (src/alg/architectures/configs.py:36-65):
//   resnet_b_l  BaseResNetActorCritic(channels = 80, num_blocks = 5)   resnet.py:24-81: conv_in + 5 residual blocks
//   cnn_b_s     BaseCnnActorCritic(channels = [56] * 4)                cnn.py:7-29: conv3x3 + BatchNorm + ReLU, stacked
//   cnn_b_l     BaseCnnActorCritic(channels = [96] * 8)
// followed in every case by the 1x1 convolutions that open the policy (C -> 2) and value (C -> 1) heads.  Eval-mode
// BatchNorm is folded into the convolutions on the host.  mnk_resnet_rows.cu / mnk_resnet.cu are specialised for the 32
// channels of resnet_b_s (the default network); this kernel is the same implicit GEMM for C = 64 (56 zero-padded), 80, 96:
//     D[pixel][c_out] = sum over (tap, c_in) A[pixel + off(tap)][c_in] * W[tap][c_out][c_in]
//   * a CTA owns floor(rows / rs) consecutive envs on `BM` UMMA M-blocks of 128 pixel rows; an env's board lies on pixel
//     rows with row stride n + 1 and n + 2 zero rows before the next env (rs = m (n+1) + n + 2), so a 3x3 tap is a constant
//     row offset and out-of-board neighbours read zeros;
//   * activations stay in shared memory for all layers, 16-bit, in the canonical K-major no-swizzle UMMA layout
//     [k-chunk = 8 channels][row][16 B]: a tap is a different start address in the descriptor.  Two ping-pong buffers
//     (A1 = the running feature map / block input, A0 = scratch);
//   * one tcgen05.mma is M = 128, N = C, K = 16: C / 16 k-steps per tap and M-block, fp32 accumulators in TMEM
//     (BM x C columns).  With N >= 64 the instruction reads 4 KB (A) + C x 32 B (B) for 2 x 128 x C x 16 flops: unlike the
//     32-channel tower it is not dominated by the operand read;
//   * a layer's weights are 9 x C x C x 2 B (115 KB at C = 80) -- too large to sit next to the activations -- so they
//     stream TAP BY TAP through a three-slot ring of 1-D TMA bulk copies issued by a producer warp; a slot is handed
//     back by the tcgen05.commit of the MMAs that read it;
//   * 16 epilogue warps (two groups on alternate M-blocks x TMEM lane quarter x channel half; the epilogue is latency-
//     bound, 8 warps took 40 % of the layer time): bias (+ skip) + ReLU -> 16-bit back to shared memory, guard rows as
//     zeros; the last layer applies the 1x1 head convolutions from the fp32 registers.
// Envs per CTA at 9x9: 5 (C <= 80, four M-blocks) or 3 (C = 96, three M-blocks: shared memory).
#include "mnk_dispatch.cuh"
#include "mnk_umma.cuh"

namespace ct {
using namespace mnk_umma;
constexpr int kMargin = 24;                   // zero rows before / after the tile (>= n + 2, i.e. n <= 22)
constexpr int kTaps = 9;
constexpr int kWtsSlots = 3;
constexpr int kEpiGroups = 2;                 // epilogue warp groups: group g takes M-blocks g, g + 2, ...
constexpr int kEpiWarps = 8 * kEpiGroups;     // per group: TMEM lane quarter = warp & 3, channel half = (warp >> 2) & 1
constexpr int kMmaWarp = kEpiWarps;           // warp 16: MMA issue
constexpr int kTmaWarp = kMmaWarp + 1;        // warp 17: weight taps
constexpr int kThreads = 32 * (kTmaWarp + 1);
constexpr int kLayerBarrier = 1;              // named barrier: epilogue warps + MMA warp at a layer boundary
constexpr int kHeadBarrier0 = 2;              // named barriers 2, 3: the 8 warps of an epilogue group (head partial sums)
constexpr int kTmemCols = 512;

template <int C, int BM>
struct Cfg {
    static constexpr int kChunks = C / 8;
    static constexpr int kRows = 128 * BM;
    static constexpr int kBufRows = kRows + 2 * kMargin;
    static constexpr int kActBytes = kChunks * kBufRows * 16;
    static constexpr int kTapBytes = kChunks * C * 16;          // [k-chunk][c_out][8 c_in]
    static constexpr int kHalf = C / 2;                          // channels per epilogue warp half
    static_assert(C % 16 == 0 && C <= 128 && BM * C <= kTmemCols, "tile shape");
};

template <int C, int BM>
struct Smem {
    alignas(128) unsigned char act[2][Cfg<C, BM>::kActBytes];
    alignas(128) unsigned char wts[kWtsSlots][Cfg<C, BM>::kTapBytes];
    alignas(16) float bias[2][C];             // this layer's folded bias (double-buffered; a TMA copy per layer)
    alignas(16) float head_w[3][C];
    float head_b[4];
    float head_part[128 * BM][3];             // last layer: partial head dot products of the upper channel half
    alignas(8) unsigned long long full_bar[kWtsSlots];    // tap weights landed
    unsigned long long free_bar[kWtsSlots];               // the MMAs that read the slot have completed
    unsigned long long mma_bar;                           // all MMAs of the layer have completed
    unsigned long long bias_bar[2];                       // the layer's bias landed
    unsigned int tmem_base;
};

struct Params {
    int m, n, words, layers;
    int residual;                     // 1: layer 0 = conv_in, then (conv1, conv2 + skip) pairs; 0: a plain conv stack
    long long num_envs;
    int spc, pw, rs;                  // envs per CTA, pixel-row stride n + 1, env stride m * pw + pw + 1
    const u64* bits;
    const u8* swap;
    const unsigned char* weights;     // op16 [layers][9 taps][C/8 k-chunks][C c_out][8 c_in]
    const float* bias;                // f32 [layers][C]
    const float* head_w;              // f32 [3][C]
    const float* head_b;              // f32 [3]
    float* policy_feat;
    float* value_feat;
    int* error;
};

MNK_DEV void tmem_ld8(u32 taddr, u32 (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
MNK_DEV void tmem_wait8(u32 (&v)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
                 :
                 : "memory");
}

template <int C, int BM>
__global__ void __launch_bounds__(kThreads, 1) conv_tower_kernel(Params p) {
    using K = Cfg<C, BM>;
    using S = Smem<C, BM>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    S& sm = *reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long env0 = (long long)blockIdx.x * p.spc;
    const int cells = p.m * p.n;
    const int envs_here = (int)min((long long)p.spc, p.num_envs - env0);
    constexpr u32 kIdesc = umma_idesc_bf16(C);
    const int total_taps = p.layers * kTaps;

    // ---- one-time setup -----------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kWtsSlots; ++i) {
            mbar_init(&sm.full_bar[i], 1);
            mbar_init(&sm.free_bar[i], 1);
        }
        mbar_init(&sm.mma_bar, 1);
        mbar_init(&sm.bias_bar[0], 1);
        mbar_init(&sm.bias_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // zero what no epilogue writes before it is read: k-chunks 0-1 of buffer 0 (the input layer's operand; the decode
        // below sets the stones) and the margins of every other plane; then the head weights
        const uint4 zero = make_uint4(0, 0, 0, 0);
        uint4* a0 = reinterpret_cast<uint4*>(&sm.act[0][0]);
        for (int i = tid; i < 2 * K::kBufRows; i += kThreads) a0[i] = zero;
        for (int i = tid; i < (2 * K::kChunks - 2) * 2 * kMargin; i += kThreads) {
            const int plane = 2 + i / (2 * kMargin), r = i % (2 * kMargin);
            a0[plane * K::kBufRows + (r < kMargin ? r : K::kRows + r)] = zero;
        }
        for (int i = tid; i < 3 * C; i += kThreads) (&sm.head_w[0][0])[i] = p.head_w[i];
        if (tid < 3) sm.head_b[tid] = p.head_b[tid];
    }
    __syncthreads();
    for (int idx = tid; idx < envs_here * cells; idx += kThreads) {     // the two canonical planes -> channels 0, 1
        const int s = idx / cells, cell = idx - s * cells;
        const int r = cell / p.n, bit = cell + r;      // bit index in the guard-strided bitboard == row offset in the tile
        const long long e = env0 + s;
        const u64 wb = p.bits[(size_t)(bit >> 6) * p.num_envs + e];
        const u64 ww = p.bits[(size_t)(p.words + (bit >> 6)) * p.num_envs + e];
        const bool sw = p.swap != nullptr && p.swap[e] != 0;
        const u32 black = (u32)(wb >> (bit & 63)) & 1u, white = (u32)(ww >> (bit & 63)) & 1u;
        const u32 me = sw ? white : black, enemy = sw ? black : white;
        reinterpret_cast<uint4*>(&sm.act[0][0])[kMargin + s * p.rs + bit] = make_uint4(me * kActOne | (enemy * kActOne) << 16, 0, 0, 0);
    }
    const int quarter = warp & 3, half = (warp >> 2) & 1, grp = warp >> 3;
    u32 valid_bits = 0;     // which of this thread's BM pixel rows are board cells: fixed for all layers
    if (warp < kEpiWarps) {
        for (int j = 0; j < BM; ++j) {
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

    if (warp == kTmaWarp) {
        // ================= weight producer: one tap (C x C x 2 B) per ring slot, three taps ahead of the MMAs ========
        for (int t = 0; t < total_taps; ++t) {
            const int slot = t % kWtsSlots;
            if (t >= kWtsSlots)
                ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.free_bar[slot], (u32)(t / kWtsSlots - 1) & 1u)) != 0;
            if (elect_one()) {
                mbar_expect_tx(&sm.full_bar[slot], K::kTapBytes);
                tma_bulk_g2s(&sm.wts[slot][0], p.weights + (size_t)t * K::kTapBytes, K::kTapBytes, &sm.full_bar[slot]);
                // The layer's bias rides along, into the buffer the epilogue of layer L - 2 read last.  (The epilogue first
                // read the bias with __ldg inside its channel loop: 10 dependent L1 / L2 round trips per thread and layer,
                // 57 % of the layer time with the tensor pipe idle -- ncu source page, profiles/README.md.)  At tap
                // kWtsSlots of layer L the wait above has seen MMAs of THIS layer complete, so the epilogue of layer L - 1
                // -- and with it every reader of this buffer -- is done.
                const int L = t / kTaps;
                if (t - L * kTaps == (L == 0 ? 0 : kWtsSlots)) {
                    mbar_expect_tx(&sm.bias_bar[L & 1], C * 4);
                    tma_bulk_g2s(&sm.bias[L & 1][0], p.bias + (size_t)L * C, C * 4, &sm.bias_bar[L & 1]);
                }
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issue: warp-uniform control flow, one elected lane issues ==========================
        int t = 0;
        for (int L = 0; L < p.layers; ++L) {
            const int in_buf = (L & 1) ? 1 : 0;
            const u64 a_d0 = umma_desc(smem_u32(&sm.act[in_buf][0]) + kMargin * 16, K::kBufRows * 16, 128);
            const u32 a_lo0 = (u32)a_d0, a_hi = (u32)(a_d0 >> 32);
            const int ksteps = (L == 0) ? 1 : C / 16;       // the input layer has 2 real channels: one K = 16 step
#pragma unroll 1
            for (int tap = 0; tap < kTaps; ++tap, ++t) {
                const int slot = t % kWtsSlots;
                ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.full_bar[slot], (u32)(t / kWtsSlots) & 1u)) != 0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u64 b_d0 = umma_desc(smem_u32(&sm.wts[slot][0]), C * 16, 128);
                const u32 b_lo0 = (u32)b_d0, b_hi = (u32)(b_d0 >> 32);
                const u32 a_tap = a_lo0 + (u32)((tap / 3 - 1) * p.pw + (tap % 3 - 1));     // rows are 16 B: row offset == field offset
                if (elect_one()) {
                    for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
                        for (int jj = 0; jj < BM; ++jj)
                            umma_bf16_lohi(tmem_base + (u32)(C * jj), a_tap + (u32)(2 * ks * K::kBufRows + 128 * jj), a_hi,
                                           b_lo0 + (u32)(2 * ks * C), b_hi, kIdesc, (tap | ks) != 0);
                    }
                    umma_commit(&sm.free_bar[slot]);          // the slot may be refilled once these MMAs have read it
                    if (tap == kTaps - 1) umma_commit(&sm.mma_bar);
                }
                __syncwarp();
            }
            // layer boundary: the next layer's operand is what the epilogue warps are about to write
            asm volatile("bar.sync %0, %1;" ::"r"(kLayerBarrier), "r"(32 * (kEpiWarps + 1)) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    } else {
        // ================= epilogue: bias (+ skip) + ReLU -> operand of the next layer / head features ============
        for (int L = 0; L < p.layers; ++L) {
            const int in_buf = (L & 1) ? 1 : 0, out_buf = in_buf ^ 1;
            const bool skip = p.residual != 0 && (L >= 2) && ((L & 1) == 0);
            const bool last = (L == p.layers - 1);
            const float* bias = &sm.bias[L & 1][K::kHalf * half];
            ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.bias_bar[L & 1], (u32)(L >> 1) & 1u)) != 0;
            ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.mma_bar, (u32)L & 1u)) != 0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = grp; j < BM; j += kEpiGroups) {
                const int i = 128 * j + quarter * 32 + lane;        // pixel row of this thread
                const bool valid = (valid_bits >> j) & 1u;
                const u32 keep = valid ? 0xFFFFFFFFu : 0u;
                const u32 taddr = tmem_base + ((u32)(quarter * 32) << 16) + (u32)(C * j + K::kHalf * half);
                uint4* out_row = reinterpret_cast<uint4*>(&sm.act[out_buf][0]) + (size_t)(K::kHalf / 8 * half) * K::kBufRows + (kMargin + i);
                float h0 = 0.f, h1 = 0.f, h2 = 0.f;
                u32 acc[2][8];
                tmem_ld8(taddr, acc[0]);
#pragma unroll
                for (int kc = 0; kc < K::kHalf / 8; ++kc) {          // 8 channels at a time; the next load is in flight
                    tmem_wait8(acc[kc & 1]);
                    if (kc + 1 < K::kHalf / 8) tmem_ld8(taddr + 8 * (kc + 1), acc[(kc + 1) & 1]);
                    const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * kc);
                    const float4 b1 = *(reinterpret_cast<const float4*>(bias + 8 * kc) + 1);
                    float v[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) v[ch] += __uint_as_float(acc[kc & 1][ch]);
                    uint4* unit = out_row + (size_t)kc * K::kBufRows;
                    if (skip) {
                        const uint4 rsd = *unit;
                        const u32 w[4] = {rsd.x, rsd.y, rsd.z, rsd.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 sk = act_unpack2(w[h]);
                            v[2 * h] += sk.x;
                            v[2 * h + 1] += sk.y;
                        }
                    }
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) v[ch] = fmaxf(v[ch], 0.0f);
                    if (!last) {
                        u32 w[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) w[h] = act_pack2(v[2 * h], v[2 * h + 1]) & keep;
                        *unit = make_uint4(w[0], w[1], w[2], w[3]);
                    } else {
                        const int c0 = K::kHalf * half + 8 * kc;
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch) {
                            h0 = fmaf(v[ch], sm.head_w[0][c0 + ch], h0);
                            h1 = fmaf(v[ch], sm.head_w[1][c0 + ch], h1);
                            h2 = fmaf(v[ch], sm.head_w[2][c0 + ch], h2);
                        }
                    }
                }
                if (last) {   // the 1x1 convolutions that open the two heads: the two channel halves meet in shared memory
                    float* part = sm.head_part[i];
                    if (half == 1) { part[0] = h0; part[1] = h1; part[2] = h2; }
                    asm volatile("bar.sync %0, %1;" ::"r"(kHeadBarrier0 + grp), "r"(256) : "memory");
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
                }
            }
            // stores -> visible to the tensor core's (async proxy) reads; TMEM reads done -> accumulators reusable
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, %1;" ::"r"(kLayerBarrier), "r"(32 * (kEpiWarps + 1)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && p.error != nullptr) atomicMax(p.error, 4);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

template <int C, int BM>
static int launch(Params& p, cudaStream_t s) {
    p.spc = Cfg<C, BM>::kRows / p.rs;
    if (p.spc < 1) return MNK_ERR_GEOM;
    const size_t smem = sizeof(Smem<C, BM>) + 128;
    if (smem > 227 * 1024) return MNK_ERR_GEOM;
    static std::atomic<size_t> granted[kMaxDevices];
    if (int rc = mnk_optin_smem(conv_tower_kernel<C, BM>, smem, granted)) return rc;
    const unsigned grid = (unsigned)((p.num_envs + p.spc - 1) / p.spc);
    conv_tower_kernel<C, BM><<<grid, kThreads, smem, s>>>(p);
    return mnk_launch_status();
}
}  // namespace ct

extern "C" int mnk_conv_tower(const mnk_state_t* st, const uint8_t* swap, int32_t channels, int32_t layers, int32_t residual,
                              const void* weights, const float* bias, const float* head_w, const float* head_b,
                              float* policy_feat, float* value_feat, int32_t* error, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (!weights || !bias || !head_w || !head_b || !policy_feat || !value_feat) return MNK_ERR_NULL;
    if (layers < 1 || layers > 32 || (residual && (layers & 1) == 0)) return MNK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(bias)) & 15u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    ct::Params p;
    p.m = st->m; p.n = st->n; p.words = st->words; p.layers = layers; p.residual = residual;
    p.num_envs = st->num_envs;
    p.pw = st->n + 1;
    p.rs = st->m * p.pw + p.pw + 1;
    if (p.pw + 1 > ct::kMargin) return MNK_ERR_GEOM;
    p.bits = reinterpret_cast<const u64*>(st->bits);
    p.swap = swap; p.weights = static_cast<const unsigned char*>(weights); p.bias = bias;
    p.head_w = head_w; p.head_b = head_b; p.policy_feat = policy_feat; p.value_feat = value_feat; p.error = error;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (channels) {
        case 64: return ct::launch<64, 4>(p, s);
        case 80: return ct::launch<80, 4>(p, s);
        case 96: return ct::launch<96, 3>(p, s);
        default: return MNK_ERR_ARG;
    }
}
