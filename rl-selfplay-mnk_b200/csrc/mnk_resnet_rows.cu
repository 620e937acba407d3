// mnk_resnet_rows.cu -- the residual tower on tcgen05, "board row per M-block" formulation (boards with m <= 10 rows).
//
// Same network and numerics as mnk_resnet.cu (reference: src/alg/architectures/resnet.py:8-95, "resnet_b_s" of
// configs.py:28-35; eval-mode BatchNorm folded; bf16 operands, fp32 accumulation); different mapping onto the tensor
// core.  mnk_resnet.cu issues one N = 32 MMA per 3x3 tap, so the activation tile is re-read from shared memory nine
// times per layer and the kernel runs at the speed of that read (40 cycles per MMA for 16 cycles of math; the
// timeline in profiles/README.md).  Here the three VERTICAL taps are fused into the MMA's N dimension:
//
//   * a CTA owns E = floor(128 / (n+1)) envs side by side; UMMA M-block b holds BOARD ROW b of all of them: lane
//     position p = s*(n+1) + c (env s, column c; one guard lane per env, unused lanes zero).  A horizontal tap is
//     a one-row shift of the shared-memory operand (a different descriptor start address, as before); a vertical
//     tap is the SAME lane of the neighbouring M-block;
//   * block b issues 3 (kx) x 2 (k-steps) MMAs of M = 128, N = 96, K = 16:
//         Q_b[p][ky*32 + co] = sum_{kx, ci} W[ky][kx][co][ci] * X_b[p + kx - 1][ci]
//     i.e. the contribution of input row b to the output rows b+1, b, b-1; the operand tile is read 3x, not 9x;
//   * the epilogue thread that owns (row r, lane p) adds three TMEM slices of ITS OWN lane,
//         out_r = Q_{r-1}[ky=0] + Q_r[ky=1] + Q_{r+1}[ky=2],
//     then bias (+ skip), ReLU, bf16 store: no shuffles, no neighbour exchange, no guard rows between board rows;
//   * the accumulators live in a ring of five 96-column TMEM slots.  The MMA warp runs up to four blocks ahead of
//     the epilogue, and because block b of layer L+1 needs only ROW b of layer L's output, the pipeline rolls
//     across layer boundaries: there is no per-layer barrier.  MMA block g waits for epilogue step g-4 (its TMEM
//     slot is free and, steps completing in order, its input row is written); epilogue step e waits for the
//     commit of MMA block e+1 (Q_{r+1}).  Per-layer weights stream through a three-slot TMA ring;
//   * hand-overs avoid shared memory.  While MMAs execute, the tensor core's operand reads own the shared-memory
//     pipe: a poll of a flag or of an mbarrier -- even a completed one -- returns ~300 cycles late (measured with
//     the -DMNK_TIMELINE build, profiles/README.md).  So "step e done" is a hardware named barrier (the 8 warps of
//     the step's epilogue set bar.arrive, the MMA warp bar.sync), and a commit-watcher warp waits on the MMA
//     mbarriers in step order and releases each step's epilogue set through another named barrier;
//   * 18 warps: two sets of 8 epilogue warps (TMEM lane quarter x channel half) take alternate steps, the MMA /
//     TMA warp, the watcher.  One CTA per SM (205 KB of shared memory at 9x9, all 512 TMEM columns).
//
// Weight layout for this kernel: bf16 [layer][kx 3][k-chunk 4][ky*32 + c_out][8 c_in] (mnk_b200/resnet.py
// arranges both layouts from the same folded parameters).
#include "mnk_dispatch.cuh"
#include "mnk_umma.cuh"


namespace rr {
using namespace mnk_umma;
constexpr int kC = 32;                        // tower channels
constexpr int kChunks = kC / 8;               // 16-byte k-chunks per pixel row
constexpr int kN = 3 * kC;                    // MMA N: (ky, c_out)
constexpr int kMaxBoardRows = 10;             // shared memory: 2 buffers x 4 k-chunks x (m*128 + 16) rows x 16 B
constexpr int kMinBoardRows = 3;
constexpr int kPad = 8;                       // zero rows before / after each activation plane (a tap shifts by one row)
constexpr int kSlots = 5;                     // TMEM ring: 5 x 96 columns
constexpr int kTmemCols = 512;
constexpr int kLead = 4;                      // MMA blocks in flight ahead of the epilogue
// Named-barrier ids are per epilogue set: step e uses index e % (2 * sets).  Two ids per set suffice and no id is shared
// between the sets -- with ids shared across sets (the first version used e % 5 like the TMEM ring) a warp of one set can
// complete a barrier phase in place of a late warp of the other set at a layer boundary; see mnk_resnet_train.cu.
constexpr int kTokenBarrier0 = 8;             // named barriers 8..11: "the blocks step e needs are committed" (id 8 + e % 4)
constexpr int kStepBarrier0 = 3;              // named barriers 3..6: "epilogue step e done" (id 3 + e % 4); 13-14: head exchange of a set
constexpr int kWtsSlots = 3;
constexpr int kLayerWeightBytes = 3 * kChunks * kN * 16;   // 18,432
#ifndef MNK_EPI_SETS
#define MNK_EPI_SETS 2
#endif
constexpr int kEpiSets = MNK_EPI_SETS;
constexpr int kBarIds = 2 * kEpiSets;                   // two sets of 8 epilogue warps take alternate steps (the step is latency-bound)
constexpr int kSetWarps = 8;                  // per set: TMEM lane quarter = warp & 3, channel half = (warp >> 2) & 1
constexpr int kMmaWarp = kEpiSets * kSetWarps;
constexpr int kWatchWarp = kMmaWarp + 1;      // turns MMA commits (mbarriers) into named-barrier tokens for the epilogue sets
constexpr int kThreads = 32 * (kWatchWarp + 1);
constexpr u32 kIdesc = umma_idesc_bf16(kN);
static_assert(kEpiSets <= kMinBoardRows && kEpiSets < kSlots, "a set's consecutive steps lie in the same or the next layer");
static_assert(kEpiSets == 2, "named-barrier ids: 3..6 step, 8..11 token, 13..14 head exchange");

struct Smem {
    alignas(128) unsigned char wts[kWtsSlots][kLayerWeightBytes];
    alignas(16) float head_w[3][kC];
    float head_b[4];
    float head_part[kEpiSets][128][3];        // last layer: partial head dot products of the upper channel half
    alignas(8) unsigned long long mma_bar[kSlots];   // MMA block g committed          (slot g % 5)
    alignas(8) unsigned long long wts_bar[kWtsSlots];
    unsigned int tmem_base;
    alignas(128) unsigned char act[1];        // [2 buffers][4 k-chunks][m*128 + 2*kPad rows][16 B], sized at launch
};

struct Params {
    int m, n, words, layers;          // layers = 1 + 2*blocks
    long long num_envs;
    int epc;                          // envs per CTA = 128 / (n+1)
    int pw;                           // lanes per env = n+1
    const u64* bits;                  // u64[2][words][num_envs]
    const uint8_t* swap;              // u8[num_envs] or null
    const unsigned char* weights;     // bf16 [layers][3][4][96][8]
    const float* bias;                // f32 [layers][32]   (BN folded)
    const float* head_w;              // f32 [3][32]: policy ch0, policy ch1, value 1x1 conv
    const float* head_b;              // f32 [3]
    float* policy_feat;               // f32 [num_envs][2*cells]
    float* value_feat;                // f32 [num_envs][cells]
    int* error;
};

MNK_DEV void tmem_ld16x3_issue(u32 t0, u32 t1, u32 t2, u32 (&a)[16], u32 (&b)[16], u32 (&c)[16]) {
    tmem_ld16_issue(t0, a);
    tmem_ld16_issue(t1, b);
    tmem_ld16_issue(t2, c);
}

__global__ void __launch_bounds__(kThreads, 1) resnet_tower_rows_kernel(Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = p.m, cells = p.m * p.n;
    const int buf_rows = m * 128 + 2 * kPad;                 // rows per k-chunk plane
    const int plane16 = buf_rows;                            // plane stride in 16-byte units
    const long long env0 = (long long)blockIdx.x * p.epc;
    const int envs_here = (int)min((long long)p.epc, p.num_envs - env0);
    uint4* const act = reinterpret_cast<uint4*>(&sm.act[0]); // [buffer*4 + chunk][row]

    // ---- one-time setup -----------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&sm.mma_bar[i], 1);
        }
        for (int i = 0; i < kWtsSlots; ++i) mbar_init(&sm.wts_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.wts_bar[0], kLayerWeightBytes);   // first layer's weights, in flight during the setup
        tma_bulk_g2s(&sm.wts[0][0], p.weights, kLayerWeightBytes, &sm.wts_bar[0]);
    }
    if (warp == kMmaWarp) {   // TMEM allocation is a warp-wide operation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // zero: k-chunks 0-1 of buffer 0 (the input layer's operand; the decode below sets the stones) and the pad rows of
        // the other six planes -- every epilogue step rewrites all 128 lanes of its row, zeros included
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < 2 * plane16; i += kThreads) act[i] = zero;
        for (int i = tid; i < 6 * 2 * kPad; i += kThreads) {
            const int plane = 2 + i / (2 * kPad), r = i % (2 * kPad);
            act[plane * plane16 + (r < kPad ? r : m * 128 + r)] = zero;
        }
        for (int i = tid; i < 3 * kC; i += kThreads) (&sm.head_w[0][0])[i] = p.head_w[i];
        if (tid < 3) sm.head_b[tid] = p.head_b[tid];
    }
    __syncthreads();
    // input: the two canonical planes of each env into channels 0,1 (k-chunk 0 of buffer 0)
    for (int idx = tid; idx < envs_here * cells; idx += kThreads) {
        const int s = idx / cells, cell = idx - s * cells;
        const int r = cell / p.n, c = cell - r * p.n, bit = cell + r;   // guard-strided bit index of the packed boards
        const long long e = env0 + s;
        const u64 wb = p.bits[(size_t)(bit >> 6) * p.num_envs + e];
        const u64 ww = p.bits[(size_t)(p.words + (bit >> 6)) * p.num_envs + e];
        const bool sw = p.swap != nullptr && p.swap[e] != 0;
        const u32 black = (u32)(wb >> (bit & 63)) & 1u, white = (u32)(ww >> (bit & 63)) & 1u;
        const u32 me = sw ? white : black, enemy = sw ? black : white;
        act[kPad + r * 128 + s * p.pw + c] = make_uint4(me * kActOne | (enemy * kActOne) << 16, 0, 0, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = sm.tmem_base;
    bool ok = true;
    const int lead = min(kLead, m);          // MMA block g waits for step g - lead (<= g - m: its input row is written too)
    const int total_steps = p.layers * m;
#ifdef MNK_ROWS_EARLY_RELEASE   // experiment, off: 0.628 ms against 0.617 ms per 32,768 envs at 9x9 (profiles/README.md)
    const bool early_release = (m == 6 || m >= 8);
#else
    const bool early_release = false;
#endif
#ifdef MNK_TIMELINE   // debug build only (tools/timeline_rows.py): cycle stamps of one mid-grid CTA into error[1..]
    const long long t_origin = clock64();
    const bool stamp = p.error != nullptr && blockIdx.x == gridDim.x / 2 && lane == 0;
#define MNK_STAMP(idx, slot) do { if (stamp && (idx) < 48) p.error[1 + 8 * (idx) + (slot)] = (int)(clock64() - t_origin); } while (0)
#else
#define MNK_STAMP(idx, slot) do { } while (0)
#endif

    if (warp == kMmaWarp) {
        // ================= MMA issue + weight streaming (warp-uniform control flow, one elected lane issues) ========
        const u32 act_lo = smem_u32(&sm.act[0]) + kPad * 16;
        int g = 0;
        for (int L = 0; L < p.layers; ++L) {
            const int in_buf = L & 1;
            const bool two_ksteps = (L != 0);   // the input layer has 2 real channels: one K=16 step
            const u64 a_d0 = umma_desc(act_lo + (u32)(in_buf * kChunks * plane16) * 16, (u32)plane16 * 16, 128);
            const u64 b_d0 = umma_desc(smem_u32(&sm.wts[L % kWtsSlots][0]), kN * 16, 128);
            const u32 a_lo0 = (u32)a_d0, a_hi = (u32)(a_d0 >> 32);   // low word: start-address field (bits 0-13) + LBO (bits 16-29)
            const u32 b_lo0 = (u32)b_d0, b_hi = (u32)(b_d0 >> 32);
            for (int b = 0; b < m; ++b, ++g) {
                // TMEM slot free (steps g-6 .. g-4 read Q_{g-5}) and input row b written (step g-m <= g-lead)
                if (g >= lead) {
                    // Hardware named barrier, not shared memory: while MMAs run, the tensor core's operand reads own the
                    // shared-memory pipe and an LDS / mbarrier poll from this warp waits ~300 cycles behind them (timeline
                    // in profiles/README.md).  The MMA warp consumes one barrier per step, in step order.
                    asm volatile("bar.sync %0, %1;" ::"r"(kStepBarrier0 + (g - lead) % kBarIds), "r"(32 * (kSetWarps + 1)) : "memory");
                }
                MNK_STAMP(g, 5);   // epilogue waits passed
                if (b == 0) {
                    // next layer's weights: its ring slot was last read by layer L-2, whose MMAs completed before
                    // the epilogue step waited for above
                    if (L + 1 < p.layers && elect_one()) {
                        mbar_expect_tx(&sm.wts_bar[(L + 1) % kWtsSlots], kLayerWeightBytes);
                        tma_bulk_g2s(&sm.wts[(L + 1) % kWtsSlots][0], p.weights + (size_t)(L + 1) * kLayerWeightBytes,
                                     kLayerWeightBytes, &sm.wts_bar[(L + 1) % kWtsSlots]);
                    }
                    ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.wts_bar[L % kWtsSlots], (u32)(L / kWtsSlots) & 1u)) != 0;
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                MNK_STAMP(g, 0);   // MMA warp: block g may be issued
                const int slot = g % kSlots;
                const u32 d_tmem = tmem_base + (u32)(slot * kN);
                const u32 a_row = a_lo0 + (u32)(b * 128 - 1);            // kx = 0 reads lane p-1
                if (elect_one()) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            if (ks == 0 || two_ksteps)
                                umma_bf16_lohi(d_tmem, a_row + (u32)kx + (u32)(2 * ks) * (u32)plane16, a_hi,
                                               b_lo0 + (u32)((kx * kChunks + 2 * ks) * kN), b_hi, kIdesc, (kx | ks) != 0);
                        }
                    }
                    umma_commit(&sm.mma_bar[slot]);
                }
                __syncwarp();
                MNK_STAMP(g, 1);   // MMA warp: block g issued + committed
            }
        }
    } else if (warp == kWatchWarp) {
        // ================= commit watcher: while MMAs run, a shared-memory poll (even of a completed mbarrier) waits ~300
        // cycles behind the tensor core's operand reads; one warp pays that, in step order, and releases the epilogue
        // set of each step through a hardware barrier ===========================================================
        for (int e = 0; e < total_steps; ++e) {
            const int r = e % m;
            const int need = (r < m - 1) ? e + 1 : e;            // Q_{r+1} is the last slice row r needs
            ok = __all_sync(MNK_FULL_WARP, ok && mbar_wait(&sm.mma_bar[need % kSlots], (u32)(need / kSlots) & 1u)) != 0;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.arrive %0, %1;" ::"r"(kTokenBarrier0 + e % kBarIds), "r"(32 * (kSetWarps + 1)) : "memory");
        }
    } else {
        // ================= epilogue: one board row (128 lanes x 32 channels) per step, sets alternate steps =========
        const int quarter = warp & 3, half = (warp >> 2) & 1, set = warp / kSetWarps;
        const int pos = quarter * 32 + lane;                   // lane position in every M-block
        const int s = pos / p.pw, c = pos - s * p.pw;
        const bool valid = s < envs_here && c < p.n;
        const u32 t_lane = tmem_base + ((u32)(quarter * 32) << 16) + (u32)(16 * half);
        // This set's steps are e = set, set + kEpiSets, ...  Layer, board row and the TMEM ring position advance
        // incrementally (the divisions and the skipped iterations of a plain loop nest were a third of the
        // instructions this kernel executed: ncu opcode histogram, profiles/README.md).
        int e = set, L = 0, r = set;                 // kEpiSets <= kMinBoardRows: the first step lies in layer 0
        int col = (set % kSlots) * kN;              // TMEM column of slot e % kSlots
        int bar = set;                              // e % kBarIds: index of this step's named barriers
        bool new_layer = true, skip = false, last = false;
        float bias[16];
        uint4* layer_out = nullptr;
        while (e < total_steps) {
            if (new_layer) {
                const int out_buf = (L & 1) ^ 1;
                skip = (L >= 2) && ((L & 1) == 0);           // second conv of a residual block adds its block input
                last = (L == p.layers - 1);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + L * kC + 16 * half) + q4);
                    bias[4 * q4] = b4.x; bias[4 * q4 + 1] = b4.y; bias[4 * q4 + 2] = b4.z; bias[4 * q4 + 3] = b4.w;
                }
                layer_out = act + (size_t)(out_buf * kChunks + 2 * half) * plane16 + (kPad + pos);
            }
            {
                if ((warp % kSetWarps) == 0) MNK_STAMP(e, 6);   // step begins (before the wait)
                asm volatile("bar.sync %0, %1;" ::"r"(kTokenBarrier0 + bar), "r"(32 * (kSetWarps + 1)) : "memory");   // token from the watcher
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if ((warp % kSetWarps) == 0) MNK_STAMP(e, 2);   // epilogue warp 0: the blocks this step needs are committed
                // out_r = Q_{r-1}[ky=0] + Q_r[ky=1] + Q_{r+1}[ky=2]; a missing neighbour row re-reads Q_r and is dropped
                const bool up = r > 0, down = r < m - 1;
                const int col_up = up ? (col == 0 ? (kSlots - 1) * kN : col - kN) : col;
                const int col_down = down ? (col == (kSlots - 1) * kN ? 0 : col + kN) : col;
                u32 q0[16], q1[16], q2[16];
                tmem_ld16x3_issue(t_lane + (u32)col_up, t_lane + (u32)(col + kC), t_lane + (u32)(col_down + 2 * kC), q0, q1, q2);
                tmem_ld_wait(q0);
                tmem_ld_wait(q1);
                tmem_ld_wait(q2);
                if ((warp % kSetWarps) == 0) MNK_STAMP(e, 3);   // TMEM slices in registers
#ifdef MNK_ROWS_EARLY_RELEASE   // experiment, off: 0.628 ms against 0.617 ms per 32,768 envs at 9x9 (profiles/README.md)
                // Early release: the step's TMEM slot is free as soon as its slices are in registers.  The release also
                // tells MMA block e + lead that ITS operand row -- written by step e + lead - m -- is in shared memory; that
                // step is an earlier step of this set (complete) when lead - m is even, and for odd m it is a step of the
                // other set that completed before that set's release of step e - 3 only if it lies at or before e - 5:
                // boards with m = 6 or m >= 8 rows (early_release); the others release at the end of the step.
                if (early_release && e + lead < total_steps) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.arrive %0, %1;" ::"r"(kStepBarrier0 + bar), "r"(32 * (kSetWarps + 1)) : "memory");
                }
#endif
                float v[16];
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {
                    float a = __uint_as_float(q1[ch]) + bias[ch];
                    if (up) a += __uint_as_float(q0[ch]);
                    if (down) a += __uint_as_float(q2[ch]);
                    v[ch] = a;
                }
                uint4* out_row[2];
                out_row[0] = layer_out + r * 128;
                out_row[1] = out_row[0] + plane16;
                if (skip) {
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
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
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) v[ch] = fmaxf(v[ch], 0.0f);
                if (!last) {
                    const u32 keep = valid ? 0xFFFFFFFFu : 0u;   // guard / unused lanes are stored as zeros
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
                        u32 w[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            w[h] = act_pack2(v[kc * 8 + 2 * h], v[kc * 8 + 2 * h + 1]) & keep;
                        }
                        *out_row[kc] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    if ((warp % kSetWarps) == 0) MNK_STAMP(e, 7);   // arithmetic + stores issued
                    // stores -> visible to the tensor core's (async proxy) reads; TMEM reads done -> slot reusable
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                } else {   // last layer: the 1x1 convolutions that open the two heads, from fp32 registers
                    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 16; ++ch) {
                        h0 = fmaf(v[ch], sm.head_w[0][16 * half + ch], h0);
                        h1 = fmaf(v[ch], sm.head_w[1][16 * half + ch], h1);
                        h2 = fmaf(v[ch], sm.head_w[2][16 * half + ch], h2);
                    }
                    float* part = sm.head_part[set][pos];
                    if (half == 1) { part[0] = h0; part[1] = h1; part[2] = h2; }
                    asm volatile("bar.sync %0, 256;" ::"r"(13 + set) : "memory");    // the 8 warps of this set (ids 13-15)
                    if (half == 0 && valid) {
                        h0 += part[0] + sm.head_b[0];
                        h1 += part[1] + sm.head_b[1];
                        h2 += part[2] + sm.head_b[2];
                        const long long env = env0 + s;
                        const int cell = r * p.n + c;
                        p.policy_feat[(size_t)env * 2 * cells + cell] = h0;
                        p.policy_feat[(size_t)env * 2 * cells + cells + cell] = h1;
                        p.value_feat[(size_t)env * cells + cell] = h2;
                    }
                    asm volatile("bar.sync %0, 256;" ::"r"(13 + set) : "memory");
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (!early_release && e + lead < total_steps)   // the last `lead` steps have no consumer
                    asm volatile("bar.arrive %0, %1;" ::"r"(kStepBarrier0 + bar), "r"(32 * (kSetWarps + 1)) : "memory");
                if ((warp % kSetWarps) == 0) MNK_STAMP(e, 4);   // step done
            }
            e += kEpiSets;
            r += kEpiSets;
            new_layer = r >= m;
            if (new_layer) { r -= m; ++L; }
            col += kEpiSets * kN;
            if (col >= kSlots * kN) col -= kSlots * kN;
            bar += kEpiSets;
            if (bar >= kBarIds) bar -= kBarIds;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && p.error != nullptr) atomicExch(p.error, 1);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}
}  // namespace rr

extern "C" int mnk_resnet_tower_rows(const mnk_state_t* st, const uint8_t* swap, const void* weights_rows, const float* bias,
                                     const float* head_w, const float* head_b, int32_t blocks, float* policy_feat,
                                     float* value_feat, int32_t* error, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (!weights_rows || !bias || !head_w || !head_b || !policy_feat || !value_feat) return MNK_ERR_NULL;
    if (blocks < 1 || blocks > 8) return MNK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(weights_rows) | reinterpret_cast<uintptr_t>(bias)) & 15u) return MNK_ERR_ALIGN;
    if (st->m < rr::kMinBoardRows || st->m > rr::kMaxBoardRows) return MNK_ERR_GEOM;
    if (st->num_envs == 0) return MNK_OK;
    rr::Params p;
    p.m = st->m; p.n = st->n; p.words = st->words; p.layers = 1 + 2 * blocks;
    p.num_envs = st->num_envs;
    p.pw = st->n + 1;
    p.epc = 128 / p.pw;
    p.bits = reinterpret_cast<const u64*>(st->bits);
    p.swap = swap; p.weights = static_cast<const unsigned char*>(weights_rows); p.bias = bias;
    p.head_w = head_w; p.head_b = head_b; p.policy_feat = policy_feat; p.value_feat = value_feat; p.error = error;
    const size_t smem = sizeof(rr::Smem) + 128 + (size_t)2 * rr::kChunks * (st->m * 128 + 2 * rr::kPad) * 16;
    static std::atomic<size_t> granted[kMaxDevices];
    if (int rc = mnk_optin_smem(rr::resnet_tower_rows_kernel, smem, granted)) return rc;
    const unsigned grid = (unsigned)((st->num_envs + p.epc - 1) / p.epc);
    rr::resnet_tower_rows_kernel<<<grid, rr::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return mnk_launch_status();
}
