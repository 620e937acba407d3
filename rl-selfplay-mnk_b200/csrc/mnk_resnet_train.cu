// mnk_resnet_train.cu -- the residual tower with TRAIN-MODE BatchNorm (batch statistics), one conv layer per launch.
//
// Reference: the rollout forward of PPOAgent.learn runs the network in train mode (src/alg/ppo.py:97; the module is
// never put in eval mode there), so every BatchNorm2d of src/alg/architectures/resnet.py:9-21,27-31 normalises with
// the mean / biased variance of the CURRENT batch over (N, H, W) and updates running_mean / running_var (momentum
// 0.1, unbiased variance).  Batch statistics of layer L need layer L of EVERY env before layer L+1 of any env can
// start, so the nine layers cannot be fused into one kernel the way the eval-mode tower is (mnk_resnet_rows.cu):
//
//   launch L (0 <= L < layers), persistent, one CTA per SM, each CTA walks groups of E = 128 / (n+1) envs:
//     * operand of the group: layer 0 decodes the bitboards; layer L > 0 receives the RAW conv output z_{L-1} of the
//       previous launch (fp16, already in the UMMA operand layout) by four 1-D TMA bulk copies, prefetched one group
//       ahead into the other half of a double buffer, and transforms it IN PLACE:
//           a_{L-1} = ReLU(scale_{L-1} * z_{L-1} + shift_{L-1} [+ a_{L-3} for the second conv of a block])
//       (guard / unused lanes -> 0).  Block inputs a_0, a_2, ... are also written to HBM: they are the skip operands
//       two launches later;
//     * the MMAs are those of the board-row kernel: M-block b = board row b of the group, 3 (kx) x 2 (k-step) MMAs of
//       M = 128, N = 96 = (ky, c_out), K = 16 into a ring of five 96-column TMEM slots, a commit-watcher warp and
//       hardware named barriers for the hand-overs (see mnk_resnet_rows.cu for why);
//     * the epilogue thread of (row r, lane p) adds its three TMEM slices, accumulates per-channel sum / sum of
//       squares over the valid lanes in registers and stores z_L as fp16 (16-byte, warp-coalesced).  The conv bias is
//       left out: BatchNorm subtracts the batch mean, so it cancels (it only enters running_mean);
//     * at the end every CTA writes its 64 partial sums; the LAST CTA to finish (a device counter) adds the partials
//       of all CTAs in a fixed order in double -- deterministic for a given grid -- and produces scale_L / shift_L
//       for the next launch plus the running-statistics update.  No separate reduction launch, no atomics on floats.
//   last launch: elementwise a_last = ReLU(BN(z_last) + skip) and the two 1x1 head convolutions -> head features.
//
// HBM traffic per env at 9x9: 9 x (6 KB z written + 6 KB read) + 4 x (6 KB a written + 6 KB read) = 156 KB -- 0.8 ms
// per 32,768 envs at the measured HBM peak, next to 0.6 ms of tensor-core time: this forward is bound by both.
#include <algorithm>

#include "mnk_dispatch.cuh"
#include "mnk_umma.cuh"

namespace rt {
using namespace mnk_umma;
constexpr int kC = 32;
constexpr int kChunks = kC / 8;
constexpr int kN = 3 * kC;
constexpr int kMaxBoardRows = 13;             // shared memory: 2 group buffers of 8 KB per board row next to 18 KB of weights
constexpr int kMinBoardRows = 3;
constexpr int kPad = 8;                       // zero rows before / after each operand plane; 1 for boards above 10 rows (shared memory)
constexpr int kSlots = 5;
constexpr int kTmemCols = 512;
constexpr int kLead = 4;
// Named-barrier ids are PER EPILOGUE SET: step e uses index e % (2 * sets) = (its set, the parity of the set's step count).
// Two ids per set are enough (a set's warps pass token e before any of them reaches token e + 2*sets, and the watcher / the
// MMA warp cannot send / consume e + 2*sets before e is complete), and no id is shared between the sets.  The first version
// indexed both families by e % 5 like the TMEM ring, so steps e (one set) and e + 5 (the other set) shared an id -- and at a
// group boundary (the step of a group's last board row needs no later block, so its token follows the previous one at once)
// a warp of the faster set could reach the sync of token e + 5 while a warp of the other set had not yet reached the sync of
// token e: the hardware counts threads, not identities, so the early warp completed e's phase in the slow warp's place, read
// its TMEM slices before the commit, and the two warps stayed exchanged until one of them ran out of steps at the end of
// the CTA's last group -- a commit-watcher timeout once in ~10,000 launches (tools/soak_cfg3.py, post-mortem in
// profiles/README.md).
constexpr int kEpiSets = 2;
constexpr int kBarIds = 2 * kEpiSets;
constexpr int kStepBarrier0 = 3;              // named barriers 3..6: "epilogue step e done" (id 3 + e % 4)
constexpr int kTokenBarrier0 = 8;             // 8..11: "the blocks step e needs are committed"
constexpr int kReadyBarrier = 13;             // "the operand of this group is in shared memory"
constexpr int kLayerWeightBytes = 3 * kChunks * kN * 16;   // 18,432
constexpr int kSetWarps = 8;
constexpr int kEpiWarps = kEpiSets * kSetWarps;
constexpr int kEpiThreads = 32 * kEpiWarps;   // 512: also the transform's thread count (4 k-chunks x 128 lanes)
constexpr int kMmaWarp = kEpiWarps;
constexpr int kWatchWarp = kMmaWarp + 1;
constexpr int kThreads = 32 * (kWatchWarp + 1);
constexpr u32 kIdesc = umma_idesc_bf16(kN);

struct Smem {
    alignas(128) unsigned char wts[kLayerWeightBytes];
    alignas(16) float scale[kC];
    float shift[kC];
    alignas(8) unsigned long long mma_bar[kSlots];
    unsigned long long wts_bar;
    unsigned long long in_bar[2];
    unsigned int tmem_base;
    unsigned int is_last;
#ifdef MNK_PROGRESS
    volatile unsigned int prog[24];           // debug build: (index << 4 | stage) of every warp, dumped on a watcher timeout
#endif
    alignas(128) unsigned char act[1];        // [2 buffers][4 k-chunks][m*128 + 2*pad rows][16 B]
};

struct Params {
    int m, n, words, layer;
    int pad;                          // zero rows before / after each operand plane
    long long num_envs, groups;
    int epc, pw;
    int reverse;                      // walk the groups from the last to the first
    const u64* bits;
    const uint8_t* swap;
    const unsigned char* z_in;        // fp16 [groups][4][m][128][8]
    const unsigned char* skip_in;     // op16, same layout, or null
    unsigned char* a_out;             // op16, same layout, or null
    unsigned char* z_out;             // fp16
    const unsigned char* weights;     // this layer: op16 [3][4][96][8], unfolded conv weights
    const float* in_scale_shift;      // f32 [64] of layer L-1
    float* partials;                  // f32 [grid][64]
    const float* gamma;               // this layer's BatchNorm weight / bias, conv bias, running statistics: f32 [32] each
    const float* beta;
    const float* conv_bias;
    float* running_mean;
    float* running_var;
    float* out_scale_shift;           // f32 [64] of layer L
    float* batch_stats;               // null or f32 [64]: batch mean (conv bias included), biased variance
    unsigned int* counter;
    unsigned int* postmortem;         // u32 [16] in the scratch buffer, zeroed per forward
    float momentum, eps;
    double count;                     // num_envs * m * n
    int* error;
};

MNK_DEV void stg128(void* ptr, uint4 v) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
MNK_DEV uint4 ldg128(const void* ptr) {
    uint4 v;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
    return v;
}
MNK_DEV float2 half2_to_float2(u32 w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
MNK_DEV u32 float2_to_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const u32*>(&h);
}

MNK_DEV void tmem_ld8_issue(u32 taddr, u32 (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
MNK_DEV void tmem_ld8_wait(u32 (&a)[8], u32 (&b)[8], u32 (&c)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                 :
                 : "memory");
}

// Pipeline (per CTA, global block / step index g = group * m + board row, groups back to back with no drain):
//   MMA block g      waits for epilogue step g - lead (its TMEM slot is free AND its operand row is in shared memory)
//   epilogue step e  first produces the operand row of block e + lead (transform of the raw fp16 row the TMA delivered, or
//                    the bitboard decode), then waits for the commit of block e + 1, moves its three TMEM slices to
//                    registers, accumulates the statistics, stores z and releases step e
//   raw z of group j is fetched by TMA into buffer j & 1 at MMA block (j-1, lead-1): every MMA of group j-2 is complete there
//                    (that block waited for step (j-2, m-1)), and the first transform of group j is lead steps away
// kFirst: the input layer (operand decoded from the bitboards); kSkip: the operand adds a skip activation; kAOut: the
// operand is also written to HBM (a block input).  Compile-time, so that the per-step code carries no dead branches.
template <bool kFirst, bool kSkip, bool kAOut>
__global__ void __launch_bounds__(kThreads, 1) resnet_layer_train_kernel(Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = p.m;
    const int pad = p.pad;
    const int plane16 = m * 128 + 2 * pad;                   // rows (16-byte units) per k-chunk plane
    const int buf16 = kChunks * plane16;                     // 16-byte units per operand buffer
    uint4* const act = reinterpret_cast<uint4*>(&sm.act[0]);
    // per-warp channel sums, written after the CTA's last MMA has completed: they reuse the weight buffer
    float (*const red)[32] = reinterpret_cast<float (*)[32]>(&sm.wts[0]);
    constexpr bool first = kFirst;
    const int my_groups = (int)((p.groups - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const size_t plane_bytes = (size_t)m * 128 * 16;
    const size_t group_bytes = kChunks * plane_bytes;
    const int total_steps = my_groups * m;
    const int lead = min(kLead, m);
    // this CTA's i-th group; consecutive launches walk the groups in opposite directions, so a launch starts on what
    // the previous one wrote last (still in L2)
    auto group_of = [&](int i) -> long long {
        const long long g = (long long)blockIdx.x + (long long)i * gridDim.x;
        return p.reverse ? p.groups - 1 - g : g;
    };

    // ---- one-time setup -----------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) mbar_init(&sm.mma_bar[i], 1);
        mbar_init(&sm.wts_bar, 1);
        mbar_init(&sm.in_bar[0], 1);
        mbar_init(&sm.in_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.wts_bar, kLayerWeightBytes);
        tma_bulk_g2s(&sm.wts[0], p.weights, kLayerWeightBytes, &sm.wts_bar);
        if (!first) {   // raw z of this CTA's first two groups
            for (int i = 0; i < min(2, my_groups); ++i) {
                const unsigned char* src = p.z_in + (size_t)group_of(i) * group_bytes;
                mbar_expect_tx(&sm.in_bar[i], (u32)group_bytes);
                for (int c = 0; c < kChunks; ++c)
                    tma_bulk_g2s(act + i * buf16 + c * plane16 + pad, src + c * plane_bytes, (u32)plane_bytes, &sm.in_bar[i]);
            }
        }
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        const uint4 zero = make_uint4(0, 0, 0, 0);
        if (first) {   // k-chunks 0-1 of both buffers (the decode sets the stones of chunk 0; chunk 1 stays zero)
            for (int i = tid; i < 2 * 2 * plane16; i += kThreads) act[(i / (2 * plane16)) * buf16 + i % (2 * plane16)] = zero;
        } else {       // the pad rows of all eight planes (the bulk copies fill rows pad .. pad + m*128)
            for (int i = tid; i < 2 * kChunks * 2 * pad; i += kThreads) {
                const int plane = i / (2 * pad), r = i % (2 * pad);
                act[plane * plane16 + (r < pad ? r : m * 128 + r)] = zero;
            }
            if (tid < 2 * kC) (&sm.scale[0])[tid] = p.in_scale_shift[tid];     // scale[32] then shift[32]
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = sm.tmem_base;
    bool ok = true;
#ifdef MNK_PROGRESS
#define MNK_PROG(idx, stage) do { if (lane == 0) sm.prog[warp] = ((unsigned)(idx) << 4) | (unsigned)(stage); } while (0)
#else
#define MNK_PROG(idx, stage) do { } while (0)
#endif

    if (warp == kMmaWarp) {
        // ================= MMA issue (one elected lane) + TMA of the raw operand two groups ahead ===================
        ok = __all_sync(MNK_FULL_WARP, mbar_wait(&sm.wts_bar, 0)) != 0;
        if (!ok && p.error != nullptr) atomicMax(p.error, 0x100 | p.layer);
        const u64 b_d0 = umma_desc(smem_u32(&sm.wts[0]), kN * 16, 128);
        const u32 b_lo0 = (u32)b_d0, b_hi = (u32)(b_d0 >> 32);
        const u32 act_lo = smem_u32(&sm.act[0]) + (u32)pad * 16;
        asm volatile("bar.sync %0, %1;" ::"r"(kReadyBarrier), "r"(kEpiThreads + 32) : "memory");   // rows 0 .. lead-1 of group 0
        int g = 0;
        for (int i = 0; i < my_groups; ++i) {
            const int buf = i & 1;
            const u64 a_d0 = umma_desc(act_lo + (u32)(buf * buf16) * 16, (u32)plane16 * 16, 128);
            const u32 a_lo0 = (u32)a_d0, a_hi = (u32)(a_d0 >> 32);
            for (int b = 0; b < m; ++b, ++g) {
                MNK_PROG(g, 1);
                if (g >= lead)
                    asm volatile("bar.sync %0, %1;" ::"r"(kStepBarrier0 + (g - lead) % kBarIds), "r"(32 * (kSetWarps + 1)) : "memory");
                MNK_PROG(g, 2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (!first && b == lead - 1 && i >= 1 && i + 1 < my_groups && elect_one()) {
                    // step (i-1, m-1) is released: every MMA that read buffer (i+1) & 1 (group i-1) has completed
                    const size_t goff = (size_t)group_of(i + 1) * group_bytes;
                    uint4* dst = act + (buf ^ 1) * buf16 + pad;
                    mbar_expect_tx(&sm.in_bar[buf ^ 1], (u32)group_bytes);
                    for (int c = 0; c < kChunks; ++c)
                        tma_bulk_g2s(dst + c * plane16, p.z_in + goff + c * plane_bytes, (u32)plane_bytes, &sm.in_bar[buf ^ 1]);
                }
                __syncwarp();
                const int slot = g % kSlots;
                const u32 d_tmem = tmem_base + (u32)(slot * kN);
                const u32 a_row = a_lo0 + (u32)(b * 128 - 1);            // kx = 0 reads lane p-1
                if (elect_one()) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            if (ks == 0 || !first)   // the input layer has 2 real channels: one K = 16 step
                                umma_bf16_lohi(d_tmem, a_row + (u32)kx + (u32)(2 * ks) * (u32)plane16, a_hi,
                                               b_lo0 + (u32)((kx * kChunks + 2 * ks) * kN), b_hi, kIdesc, (kx | ks) != 0);
                        }
                    }
                    umma_commit(&sm.mma_bar[slot]);
                }
                __syncwarp();
                MNK_PROG(g, 3);
            }
        }
        MNK_PROG(g, 7);
    } else if (warp == kWatchWarp) {
        // ================= commit watcher: MMA commits (mbarriers) -> named-barrier tokens, in step order ==========
        for (int e = 0; e < total_steps; ++e) {
            const int r = e % m;
            const int need = (r < m - 1) ? e + 1 : e;
            const bool was = ok;
            MNK_PROG(e, 1);
            ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.mma_bar[need % kSlots], (u32)(need / kSlots) & 1u)) != 0;
            if (was && !ok && p.error != nullptr && lane == 0) atomicMax(p.error, 0x200 | p.layer);
            // post-mortem of the first timeout (never taken in a healthy run): which CTA / step / layer, and the five commit
            // barriers' raw words, into the scratch buffer (tools/debug_train.py prints them)
            if (was && !ok) {
                if (lane == 0 && atomicAdd(p.postmortem, 1u) == 0u) {
                    p.postmortem[1] = blockIdx.x | ((unsigned)total_steps << 16);
                    p.postmortem[2] = (unsigned)e | ((unsigned)p.layer << 16);
                    p.postmortem[3] = (unsigned)my_groups;
                    for (int q = 0; q < kSlots; ++q) {
                        const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(&sm.mma_bar[q]);
                        p.postmortem[4 + 2 * q] = (unsigned)(w & 0xFFFFFFFFull);
                        p.postmortem[5 + 2 * q] = (unsigned)(w >> 32);
                    }
#ifdef MNK_PROGRESS
                    for (int q = 0; q < 18; ++q) p.postmortem[14 + q] = sm.prog[q];
                    for (int q = 0; q < 2; ++q) {
                        const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(&sm.in_bar[q]);
                        p.postmortem[32 + 2 * q] = (unsigned)(w & 0xFFFFFFFFull);
                        p.postmortem[33 + 2 * q] = (unsigned)(w >> 32);
                    }
#endif
                    __threadfence();
                }
                __syncwarp();      // no token leaves before the dump is complete
            }
            MNK_PROG(e, 2);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.arrive %0, %1;" ::"r"(kTokenBarrier0 + e % kBarIds), "r"(32 * (kSetWarps + 1)) : "memory");
        }
    } else {
        // ================= operand rows + epilogue (two sets of 8 warps on alternate steps) ========================
        // Everything that depends only on the group (its index in HBM, how many of its envs exist, this thread's
        // validity, base pointers) is refreshed when the group changes, not every step: integer / address arithmetic was
        // half of the ~400 instructions a warp executed per step (ncu opcode histogram, profiles/README.md).
        const int quarter = warp & 3, half = (warp >> 2) & 1, set = warp / kSetWarps;
        const int pos = quarter * 32 + lane;                             // epilogue: TMEM lane
        const int e_s = pos / p.pw, e_c = pos - e_s * p.pw;
        const u32 t_lane = tmem_base + ((u32)(quarter * 32) << 16) + (u32)(16 * half);
        // operand rows: a set's 256 threads cover one row (128 lanes x 4 k-chunks): lane t_pos, k-chunks t_c0 and t_c0 + 2
        const int tis = tid & (32 * kSetWarps - 1);
        const int t_c0 = tis >> 7, t_pos = tis & 127;
        const int t_s = t_pos / p.pw, t_c = t_pos - t_s * p.pw;
        const bool t_lane_ok = t_s < p.epc && t_c < p.n;
        uint4* const unit0 = act + t_c0 * plane16 + pad + t_pos;        // this thread's unit of row 0, buffer 0, first k-chunk
        const size_t t_off = (size_t)t_c0 * plane_bytes + (size_t)t_pos * 16;    // the same unit inside a group in HBM
        const size_t e_off = (size_t)(2 * half) * plane_bytes + (size_t)pos * 16;
        // producer-side group state (group jt of the operand row being produced) and epilogue-side (group j)
        int pg = -1;
        u32 p_keep = 0;
        size_t p_goff = 0;
        long long p_env = 0;
        bool p_here = false;
        auto enter_producer_group = [&](int jt) {
            pg = jt;
            const long long G = group_of(jt);
            const int envs_here = (int)min((long long)p.epc, p.num_envs - G * p.epc);
            p_keep = (t_s < envs_here && t_c < p.n) ? 0xFFFFFFFFu : 0u;
            p_goff = (size_t)G * group_bytes + t_off;
            p_env = G * p.epc + t_s;
            p_here = t_c0 == 0 && t_lane_ok && t_s < envs_here;
            if constexpr (!kFirst) {
                // buffer jt & 1 holds the raw rows of group jt.  The vote also re-converges the warp after the spin loop:
                // the named barriers and tcgen05.ld below are .aligned (all 32 lanes must execute them together)
                const bool landed = __all_sync(MNK_FULL_WARP, mbar_wait(&sm.in_bar[jt & 1], (u32)(jt >> 1) & 1u)) != 0;
                if (!landed && ok && p.error != nullptr) atomicMax(p.error, 0x400 | p.layer);
                ok = ok && landed;
            }
        };
        // raw material of the operand row (jt, bt), fetched one own step ahead: the skip operand (two 16-byte units), or
        // for the input layer the packed stones of this lane's cell
        auto fetch = [&](int jt, int bt, uint4& f0, uint4& f1) {
            if constexpr (kFirst) {
                const long long G = group_of(jt);
                const long long en = G * p.epc + t_s;
                f0 = make_uint4(0, 0, 0, 0);
                if (t_c0 == 0 && t_lane_ok && en < p.num_envs) {
                    const int bit = bt * p.pw + t_c;
                    const u64 wb = p.bits[(size_t)(bit >> 6) * p.num_envs + en];
                    const u64 ww = p.bits[(size_t)(p.words + (bit >> 6)) * p.num_envs + en];
                    const u32 black = (u32)(wb >> (bit & 63)) & 1u, white = (u32)(ww >> (bit & 63)) & 1u;
                    const bool sw = p.swap != nullptr && p.swap[en] != 0;
                    f0.x = (sw ? white : black) * kActOne | ((sw ? black : white) * kActOne) << 16;
                }
            } else if constexpr (kSkip) {
                const unsigned char* src = p.skip_in + (size_t)group_of(jt) * group_bytes + t_off + (size_t)bt * 2048;
                f0 = ldg128(src);
                f1 = ldg128(src + 2 * plane_bytes);
            }
        };
        auto produce = [&](int jt, int bt, const uint4& f0, const uint4& f1) {
            if (jt != pg) enter_producer_group(jt);
            uint4* row = unit0 + (jt & 1) * buf16 + bt * 128;
            if constexpr (kFirst) {
                if (t_c0 == 0 && t_lane_ok) *row = f0;
                return;
            } else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    uint4* unit = row + u * 2 * plane16;
                    const uint4 zq = *unit;
                    const uint4 sk = u == 0 ? f0 : f1;
                    const float4* ss = reinterpret_cast<const float4*>(&sm.scale[(t_c0 + 2 * u) * 8]);     // shift follows scale
                    const float4 sc0 = ss[0], sc1 = ss[1], sh0 = ss[8], sh1 = ss[9];
                    const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                    const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
                    const u32 zw[4] = {zq.x, zq.y, zq.z, zq.w};
                    const u32 kw[4] = {sk.x, sk.y, sk.z, sk.w};
                    u32 ow[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float2 z2 = half2_to_float2(zw[h]);
                        float y0 = fmaf(z2.x, sc[2 * h], sh[2 * h]);
                        float y1 = fmaf(z2.y, sc[2 * h + 1], sh[2 * h + 1]);
                        if constexpr (kSkip) {
                            const float2 k2 = act_unpack2(kw[h]);
                            y0 += k2.x;
                            y1 += k2.y;
                        }
                        ow[h] = act_pack2(fmaxf(y0, 0.0f), fmaxf(y1, 0.0f)) & p_keep;     // guard / unused lanes -> 0
                    }
                    const uint4 out = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    *unit = out;
                    if constexpr (kAOut) stg128(p.a_out + p_goff + (size_t)u * 2 * plane_bytes + (size_t)bt * 2048, out);
                }
            }
        };

        // ---- prologue: operand rows 0 .. lead-1 of this CTA's first group (both sets, two rows each pass) ------------
        for (int bt = set; bt < lead; bt += kEpiSets) {
            uint4 f0 = make_uint4(0, 0, 0, 0), f1 = f0;
            fetch(0, bt, f0, f1);
            produce(0, bt, f0, f1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.arrive %0, %1;" ::"r"(kReadyBarrier), "r"(kEpiThreads + 32) : "memory");

        float s1[16], s2[16];
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) s1[ch] = s2[ch] = 0.0f;
        int e = set, j = 0, r = set;                         // this set's step: group j, board row r (m >= 3 > set)
        int jt = (set + lead) / m, bt = (set + lead) - jt * m;     // the operand row this step produces: block e + lead
        int col = (set % kSlots) * kN;
        int bar = set;                                       // e % kBarIds: index of this step's named barriers
        int eg = -1;                                         // epilogue-side group state
        bool e_valid = false;
        unsigned char* zgroup = nullptr;
        uint4 f0 = make_uint4(0, 0, 0, 0), f1 = f0;
        if (jt < my_groups) fetch(jt, bt, f0, f1);
        while (e < total_steps) {
            MNK_PROG(e, 1);
            if (jt < my_groups) {
                produce(jt, bt, f0, f1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            MNK_PROG(e, 2);
            asm volatile("bar.sync %0, %1;" ::"r"(kTokenBarrier0 + bar), "r"(32 * (kSetWarps + 1)) : "memory");
            MNK_PROG(e, 3);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool up = r > 0, down = r < m - 1;
            const int col_up = up ? (col == 0 ? (kSlots - 1) * kN : col - kN) : col;
            const int col_down = down ? (col == (kSlots - 1) * kN ? 0 : col + kN) : col;
            float v[16];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {     // 8 channels at a time: 24 registers in flight instead of 48
                u32 q0[8], q1[8], q2[8];
                tmem_ld8_issue(t_lane + (u32)(col_up + 8 * hh), q0);
                tmem_ld8_issue(t_lane + (u32)(col + kC + 8 * hh), q1);
                tmem_ld8_issue(t_lane + (u32)(col_down + 2 * kC + 8 * hh), q2);
                tmem_ld8_wait(q0, q1, q2);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    float a = __uint_as_float(q1[ch]);
                    if (up) a += __uint_as_float(q0[ch]);
                    if (down) a += __uint_as_float(q2[ch]);
                    v[8 * hh + ch] = a;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            MNK_PROG(e, 4);
            const bool arrive_now = jt < my_groups;
            // Step e is released HERE: its TMEM slices are in registers (slot free) and the operand row of block e + lead was
            // written before the token sync; the statistics and the stores of row e run under the next MMAs.  (With the
            // named-barrier ids shared between the sets -- see kBarIds -- this placement deadlocked every second forward and
            // the release at the very end of the step once in ~10,000 launches; with per-set ids both are clean over 54,000
            // launches of tools/soak_cfg3.py and this one is 5 % faster.)
            if (arrive_now)
                asm volatile("bar.arrive %0, %1;" ::"r"(kStepBarrier0 + bar), "r"(32 * (kSetWarps + 1)) : "memory");
            // raw material for this set's next step
            bt += kEpiSets;
            if (bt >= m) { bt -= m; ++jt; }
            if (jt < my_groups) fetch(jt, bt, f0, f1);
            // statistics + fp16 store of row (j, r), from registers
            if (j != eg) {
                eg = j;
                const long long G = group_of(j);
                const int envs_here = (int)min((long long)p.epc, p.num_envs - G * p.epc);
                e_valid = e_s < envs_here && e_c < p.n;
                zgroup = p.z_out + (size_t)G * group_bytes + e_off;
            }
            {
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {
                    v[ch] = e_valid ? v[ch] : 0.0f;
                    s1[ch] += v[ch];
                    s2[ch] = fmaf(v[ch], v[ch], s2[ch]);
                }
                unsigned char* zrow = zgroup + (size_t)r * 2048;
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                    u32 w[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) w[h] = float2_to_half2(v[kc * 8 + 2 * h], v[kc * 8 + 2 * h + 1]);
                    stg128(zrow + (size_t)kc * plane_bytes, make_uint4(w[0], w[1], w[2], w[3]));
                }
            }
            MNK_PROG(e, 5);
            MNK_PROG(e, 6);
            e += kEpiSets;
            r += kEpiSets;
            if (r >= m) { r -= m; ++j; }
            col += kEpiSets * kN;
            if (col >= kSlots * kN) col -= kSlots * kN;
            bar += kEpiSets;
            if (bar >= kBarIds) bar -= kBarIds;
        }
        MNK_PROG(e, 7);
        // per-warp sums of this warp's 16 channels over its 32 lanes
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s1[ch] += __shfl_xor_sync(MNK_FULL_WARP, s1[ch], off);
                s2[ch] += __shfl_xor_sync(MNK_FULL_WARP, s2[ch], off);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {
                red[warp][ch] = s1[ch];
                red[warp][16 + ch] = s2[ch];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && p.error != nullptr) atomicMax(p.error, 1);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
    // ---- batch statistics: this CTA's partial sums; the last CTA reduces all of them in a fixed order ----------
    if (tid < 64) {
        const int which = tid >> 5, ch = tid & 31, half = ch >> 4;
        float acc = 0.0f;
        for (int w = 0; w < kEpiWarps; ++w)
            if (((w >> 2) & 1) == half) acc += red[w][which * 16 + (ch & 15)];
        p.partials[(size_t)blockIdx.x * 64 + tid] = acc;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(p.counter, 1u);
        sm.is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (sm.is_last != 0u && tid < kC) {
        __threadfence();
        double sum = 0.0, sq = 0.0;
        for (unsigned int c = 0; c < gridDim.x; ++c) {
            sum += (double)__ldcg(p.partials + (size_t)c * 64 + tid);
            sq += (double)__ldcg(p.partials + (size_t)c * 64 + 32 + tid);
        }
        const double mean = sum / p.count;
        const double var = fmax(sq / p.count - mean * mean, 0.0);
        const float scale = (float)((double)p.gamma[tid] / sqrt(var + (double)p.eps));
        p.out_scale_shift[tid] = scale;
        p.out_scale_shift[32 + tid] = (float)((double)p.beta[tid] - mean * (double)scale);
        const float mean_x = (float)(mean + (double)p.conv_bias[tid]);       // the conv bias was left out of z
        const double unbiased = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
        p.running_mean[tid] = (1.0f - p.momentum) * p.running_mean[tid] + p.momentum * mean_x;
        p.running_var[tid] = (1.0f - p.momentum) * p.running_var[tid] + p.momentum * (float)unbiased;
        if (p.batch_stats != nullptr) {
            p.batch_stats[tid] = mean_x;
            p.batch_stats[32 + tid] = (float)var;
        }
        if (tid == 0) *p.counter = 0u;
    }
}

// a_last = ReLU(BN(z_last) [+ skip]) and the 1x1 convolutions that open the two heads; one thread per (group, row, lane)
struct FeatParams {
    int m, n, epc, pw;
    int reverse;             // start with the groups the last layer launch wrote last
    long long num_envs, groups;
    const unsigned char* z_in;
    const unsigned char* skip_in;
    const float* scale_shift;
    const float* head_w;     // f32 [3][32]
    const float* head_b;     // f32 [3]
    float* policy_feat;
    float* value_feat;
};

__global__ void __launch_bounds__(256) resnet_train_features_kernel(FeatParams p) {
    __shared__ float ss[64], hw[96], hb[3];
    if (threadIdx.x < 64) ss[threadIdx.x] = p.scale_shift[threadIdx.x];
    if (threadIdx.x < 96) hw[threadIdx.x] = p.head_w[threadIdx.x];
    if (threadIdx.x < 3) hb[threadIdx.x] = p.head_b[threadIdx.x];
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_group = (long long)p.m * 128;
    if (idx >= p.groups * per_group) return;
    const long long G = p.reverse ? p.groups - 1 - idx / per_group : idx / per_group;
    const int rp = (int)(idx % per_group), r = rp >> 7, pos = rp & 127;
    const int s = pos / p.pw, c = pos - s * p.pw;
    const long long env = G * p.epc + s;
    if (s >= p.epc || c >= p.n || env >= p.num_envs) return;
    const size_t plane_bytes = (size_t)p.m * 2048;
    const size_t off = (size_t)G * kChunks * plane_bytes + (size_t)rp * 16;
    float h0 = hb[0], h1 = hb[1], h2 = hb[2];
#pragma unroll
    for (int kc = 0; kc < kChunks; ++kc) {
        const uint4 zq = ldg128(p.z_in + off + kc * plane_bytes);
        uint4 sq = make_uint4(0, 0, 0, 0);
        if (p.skip_in != nullptr) sq = ldg128(p.skip_in + off + kc * plane_bytes);
        const u32 zw[4] = {zq.x, zq.y, zq.z, zq.w}, sw[4] = {sq.x, sq.y, sq.z, sq.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float2 z2 = half2_to_float2(zw[h]);
            const int ch = kc * 8 + 2 * h;
            float y0 = fmaf(z2.x, ss[ch], ss[32 + ch]), y1 = fmaf(z2.y, ss[ch + 1], ss[32 + ch + 1]);
            if (p.skip_in != nullptr) {
                const float2 k2 = act_unpack2(sw[h]);
                y0 += k2.x;
                y1 += k2.y;
            }
            y0 = fmaxf(y0, 0.0f);
            y1 = fmaxf(y1, 0.0f);
            h0 = fmaf(y0, hw[ch], fmaf(y1, hw[ch + 1], h0));
            h1 = fmaf(y0, hw[32 + ch], fmaf(y1, hw[32 + ch + 1], h1));
            h2 = fmaf(y0, hw[64 + ch], fmaf(y1, hw[64 + ch + 1], h2));
        }
    }
    const int cells = p.m * p.n, cell = r * p.n + c;
    p.policy_feat[(size_t)env * 2 * cells + cell] = h0;
    p.policy_feat[(size_t)env * 2 * cells + cells + cell] = h1;
    p.value_feat[(size_t)env * cells + cell] = h2;
}

struct Layout {
    size_t act_bytes;       // one activation array: groups * 4 * m * 128 * 16
    size_t z[2], a[2], partials, scale_shift, counter, total;
    long long groups;
    int grid;
};

static inline Layout layout_for(int m, int n, long long num_envs, int layers) {
    Layout l;
    const int epc = 128 / (n + 1);
    l.groups = (num_envs + epc - 1) / epc;
    l.grid = (int)std::min<long long>(std::max<long long>(l.groups, 1), mnk_sm_count());
    l.act_bytes = (size_t)l.groups * kChunks * m * 128 * 16;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
    l.z[0] = take(l.act_bytes); l.z[1] = take(l.act_bytes);
    l.a[0] = take(l.act_bytes); l.a[1] = take(l.act_bytes);
    l.partials = take((size_t)l.grid * 64 * sizeof(float));
    l.scale_shift = take((size_t)layers * 64 * sizeof(float));
    l.counter = take(256);
    l.total = off;
    return l;
}
}  // namespace rt

extern "C" int64_t mnk_resnet_tower_train_scratch_bytes(int32_t m, int32_t n, int64_t num_envs, int32_t blocks) {
    if (m < rt::kMinBoardRows || m > rt::kMaxBoardRows || n < 1 || n > 32 || num_envs < 0 || blocks < 0 || blocks > 8) return MNK_ERR_GEOM;
    return (int64_t)rt::layout_for(m, n, num_envs, 1 + 2 * blocks).total;
}

extern "C" int mnk_resnet_tower_train(const mnk_state_t* st, const uint8_t* swap, const void* weights_rows, const mnk_bn_train_t* bn,
                                      const float* head_w, const float* head_b, int32_t blocks, void* scratch, int64_t scratch_bytes,
                                      float* policy_feat, float* value_feat, int32_t* error, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (!weights_rows || !bn || !head_w || !head_b || !policy_feat || !value_feat || !scratch) return MNK_ERR_NULL;
    if (!bn->gamma || !bn->beta || !bn->conv_bias || !bn->running_mean || !bn->running_var) return MNK_ERR_NULL;
    if (blocks < 0 || blocks > 8) return MNK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(weights_rows) & 15u) || (reinterpret_cast<uintptr_t>(scratch) & 255u)) return MNK_ERR_ALIGN;
    if (st->m < rt::kMinBoardRows || st->m > rt::kMaxBoardRows) return MNK_ERR_GEOM;
    if (st->num_envs == 0) return MNK_OK;
    const int layers = 1 + 2 * blocks;
    const rt::Layout lay = rt::layout_for(st->m, st->n, st->num_envs, layers);
    if (scratch_bytes < (int64_t)lay.total) return MNK_ERR_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char* base = static_cast<unsigned char*>(scratch);
    float* scale_shift = reinterpret_cast<float*>(base + lay.scale_shift);
    // the CTA counter is re-armed every forward; the post-mortem words behind it (+64 B) are STICKY -- they keep the first
    // timeout since the caller zeroed the scratch buffer, so a failure inside a long captured rollout can still be read
    cudaError_t e = cudaMemsetAsync(base + lay.counter, 0, 64, s);
    if (e != cudaSuccess) return (int)e;
    const int pad = st->m <= 10 ? rt::kPad : 1;
    const size_t smem = sizeof(rt::Smem) + 128 + (size_t)2 * rt::kChunks * (st->m * 128 + 2 * pad) * 16;
    if (smem > 227 * 1024) return MNK_ERR_GEOM;
    static std::atomic<size_t> granted[5][kMaxDevices];
    if (int rc = mnk_optin_smem(rt::resnet_layer_train_kernel<true, false, false>, smem, granted[0])) return rc;
    if (int rc = mnk_optin_smem(rt::resnet_layer_train_kernel<false, true, true>, smem, granted[1])) return rc;
    if (int rc = mnk_optin_smem(rt::resnet_layer_train_kernel<false, true, false>, smem, granted[2])) return rc;
    if (int rc = mnk_optin_smem(rt::resnet_layer_train_kernel<false, false, true>, smem, granted[3])) return rc;
    if (int rc = mnk_optin_smem(rt::resnet_layer_train_kernel<false, false, false>, smem, granted[4])) return rc;
    for (int L = 0; L < layers; ++L) {
        rt::Params p;
        p.m = st->m; p.n = st->n; p.words = st->words; p.layer = L; p.pad = pad;
        p.num_envs = st->num_envs; p.groups = lay.groups;
        p.pw = st->n + 1; p.epc = 128 / p.pw;
        p.reverse = L & 1;
        p.bits = reinterpret_cast<const u64*>(st->bits); p.swap = swap;
        const int in = L - 1;                                  // the operand of this launch is a_in
        p.z_in = L > 0 ? base + lay.z[in & 1] : nullptr;
        p.skip_in = (L > 0 && in >= 2 && (in & 1) == 0) ? base + lay.a[((in - 2) / 2) & 1] : nullptr;
        p.a_out = (L > 0 && (in & 1) == 0 && in + 2 < layers) ? base + lay.a[(in / 2) & 1] : nullptr;
        p.z_out = base + lay.z[L & 1];
        p.weights = static_cast<const unsigned char*>(weights_rows) + (size_t)L * rt::kLayerWeightBytes;
        p.in_scale_shift = L > 0 ? scale_shift + (size_t)in * 64 : nullptr;
        p.partials = reinterpret_cast<float*>(base + lay.partials);
        p.gamma = bn->gamma + L * 32; p.beta = bn->beta + L * 32; p.conv_bias = bn->conv_bias + L * 32;
        p.running_mean = bn->running_mean + L * 32; p.running_var = bn->running_var + L * 32;
        p.out_scale_shift = scale_shift + (size_t)L * 64;
        p.batch_stats = bn->batch_stats ? bn->batch_stats + L * 64 : nullptr;
        p.counter = reinterpret_cast<unsigned int*>(base + lay.counter);
        p.postmortem = p.counter + 16;
        p.momentum = bn->momentum; p.eps = bn->eps;
        p.count = (double)st->num_envs * st->m * st->n;
        p.error = error;
        const bool skip = p.skip_in != nullptr, aout = p.a_out != nullptr;
        if (L == 0) rt::resnet_layer_train_kernel<true, false, false><<<lay.grid, rt::kThreads, smem, s>>>(p);
        else if (skip && aout) rt::resnet_layer_train_kernel<false, true, true><<<lay.grid, rt::kThreads, smem, s>>>(p);
        else if (skip) rt::resnet_layer_train_kernel<false, true, false><<<lay.grid, rt::kThreads, smem, s>>>(p);
        else if (aout) rt::resnet_layer_train_kernel<false, false, true><<<lay.grid, rt::kThreads, smem, s>>>(p);
        else rt::resnet_layer_train_kernel<false, false, false><<<lay.grid, rt::kThreads, smem, s>>>(p);
        if (int rc = mnk_launch_status()) return rc;
    }
    rt::FeatParams f;
    const int last = layers - 1;
    f.m = st->m; f.n = st->n; f.pw = st->n + 1; f.epc = 128 / f.pw;
    f.num_envs = st->num_envs; f.groups = lay.groups;
    f.reverse = (last & 1) ? 0 : 1;
    f.z_in = base + lay.z[last & 1];
    f.skip_in = (last >= 2 && (last & 1) == 0) ? base + lay.a[((last - 2) / 2) & 1] : nullptr;
    f.scale_shift = scale_shift + (size_t)last * 64;
    f.head_w = head_w; f.head_b = head_b; f.policy_feat = policy_feat; f.value_feat = value_feat;
    const long long threads = lay.groups * st->m * 128;
    rt::resnet_train_features_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(f);
    return mnk_launch_status();
}
