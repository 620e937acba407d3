// mnk_device.cuh -- device-side building blocks shared by every kernel of libmnk_b200.so.
//
// Bitboard geometry (see include/mnk_b200.h): a plane is NW uint64 words, bit r*(n+1)+c <=> cell
// (r, c); column n of every row is a permanently-zero guard.  All multi-word helpers below are
// written as fully unrolled, predicated loops over compile-time word counts so that the arrays stay
// in registers; with a static geometry (SGeom) every offset is a compile-time constant and the
// predicates fold away, with a dynamic geometry (DGeom) the same code runs with runtime offsets.
//
// Reference behaviour restated here (paths relative to the reference repo):
//   apply_move  <- src/env/torch_vector_mnk_env.py:60-84   (step_subset)
//   has_line    <- src/env/torch_vector_mnk_env.py:106-119 (_check_wins: three valid conv2d)
//   build_views <- src/env/torch_vector_mnk_env.py:46-53   (observe) and
//                  src/selfplay/torch_self_play_wrapper.py:99-112 (_get_canonical_obs)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mnk_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

#define MNK_FULL_WARP 0xffffffffu
#define MNK_DEV __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
template <int M, int N, int K>
struct SGeom {
    static constexpr bool kStatic = true;
    static constexpr int NW = (M * (N + 1) + 63) / 64;   // guard-strided words per plane
    static constexpr int NWD = (2 * M * N + 63) / 64;    // dense words for both planes (2*m*n bits)
    static constexpr int NWL = (M * N + 63) / 64;        // dense words for one plane / the legal set
    MNK_DEV constexpr int m() const { return M; }
    MNK_DEV constexpr int n() const { return N; }
    MNK_DEV constexpr int k() const { return K; }
    MNK_DEV constexpr int stride() const { return N + 1; }
    MNK_DEV constexpr int cells() const { return M * N; }
};

template <int NW_>
struct DGeom {
    static constexpr bool kStatic = false;
    static constexpr int NW = NW_;
    static constexpr int NWD = 2 * NW_;   // 2*m*n < 2*m*(n+1) <= 128*NW
    static constexpr int NWL = NW_;
    int m_, n_, k_;
    MNK_DEV int m() const { return m_; }
    MNK_DEV int n() const { return n_; }
    MNK_DEV int k() const { return k_; }
    MNK_DEV int stride() const { return n_ + 1; }
    MNK_DEV int cells() const { return m_ * n_; }
};

// ------------------------------------------------------------------------------------------------
// multi-word bit helpers (arrays of u64, little-endian bit order)
// ------------------------------------------------------------------------------------------------
template <int W>
MNK_DEV u64 word_at(const u64 (&a)[W], int i) {   // a[i] with a runtime index, 0 when out of range
    u64 v = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) v = (w == i) ? a[w] : v;
    return v;
}

// bits [off, off+nbits) of a, nbits <= 32
template <int W>
MNK_DEV u32 get_field(const u64 (&a)[W], int off, int nbits) {
    const int wi = off >> 6, sh = off & 63;
    u64 lo = word_at(a, wi) >> sh;
    if (sh + nbits > 64) lo |= word_at(a, wi + 1) << (64 - sh);
    return (u32)lo & (nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u));
}

// a |= val << off   (val holds <= 32 significant bits)
template <int W>
MNK_DEV void or_field(u64 (&a)[W], int off, u32 val) {
    const int wi = off >> 6, sh = off & 63;
    const u64 lo = (u64)val << sh;
    const u64 hi = sh > 32 ? ((u64)val >> (64 - sh)) : 0ull;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if (w == wi) a[w] |= lo;
        if (w == wi + 1) a[w] |= hi;
    }
}

// out = a >> s  (0 <= s < 64*W)
template <int W>
MNK_DEV void shr_words(const u64 (&a)[W], int s, u64 (&out)[W]) {
    const int q = s >> 6, r = s & 63;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const u64 lo = word_at(a, w + q);
        const u64 hi = word_at(a, w + q + 1);
        out[w] = r ? ((lo >> r) | (hi << (64 - r))) : lo;
    }
}

template <int W>
MNK_DEV bool any_bit(const u64 (&a)[W]) {
    u64 v = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) v |= a[w];
    return v != 0;
}

// position of the j-th (0-based) set bit of x; requires j < popc(x)
MNK_DEV int nth_set_bit64(u64 x, int j) {
    int pos = 0;
    u32 v = (u32)x;
    int c = __popc(v);
    if (j >= c) { j -= c; pos = 32; v = (u32)(x >> 32); }
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const u32 lowmask = (1u << half) - 1u;
        c = __popc(v & lowmask);
        if (j >= c) { j -= c; pos += half; v >>= half; }
        v &= lowmask;
    }
    return pos;
}

// ------------------------------------------------------------------------------------------------
// k-in-a-row: shift-and-AND run-length doubling in the four directions {1, W, W+1, W-1}.
// run_L has bit i set iff bits i, i+d, ..., i+(L-1)d are all set; run_{L+r} = run_L & (run_L >> r*d)
// for r <= L.  The guard column makes wrapped lines impossible, which is exactly what the
// reference's *valid* convolutions compute (torch_vector_mnk_env.py:110-112); >= k in a row
// (overlines) wins because some window of k inside it is full.
// ------------------------------------------------------------------------------------------------
template <class G>
MNK_DEV bool has_line(const G& g, const u64 (&b)[G::NW]) {
    const int W = g.stride();
    const int k = g.k();
    bool found = false;
#pragma unroll
    for (int dir = 0; dir < 4; ++dir) {
        const int d = dir == 0 ? 1 : dir == 1 ? W : dir == 2 ? W + 1 : W - 1;
        u64 run[G::NW], tmp[G::NW];
#pragma unroll
        for (int w = 0; w < G::NW; ++w) run[w] = b[w];
        int len = 1;
#pragma unroll 1
        for (; 2 * len <= k; len *= 2) {
            shr_words(run, len * d, tmp);
#pragma unroll
            for (int w = 0; w < G::NW; ++w) run[w] &= tmp[w];
        }
        if (len < k) {
            shr_words(run, (k - len) * d, tmp);
#pragma unroll
            for (int w = 0; w < G::NW; ++w) run[w] &= tmp[w];
        }
        found |= any_bit(run);
    }
    return found;
}

// static-geometry specialisation of the doubling loop: every shift distance is a constant
template <int M, int N, int K>
MNK_DEV bool has_line(const SGeom<M, N, K>&, const u64 (&b)[SGeom<M, N, K>::NW]) {
    constexpr int NW = SGeom<M, N, K>::NW;
    constexpr int W = N + 1;
    u64 acc = 0;
#pragma unroll
    for (int dir = 0; dir < 4; ++dir) {
        const int d = dir == 0 ? 1 : dir == 1 ? W : dir == 2 ? W + 1 : W - 1;
        u64 run[NW], tmp[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) run[w] = b[w];
        int len = 1;
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            if (2 * len <= K) {
                shr_words(run, len * d, tmp);
#pragma unroll
                for (int w = 0; w < NW; ++w) run[w] &= tmp[w];
                len *= 2;
            }
        }
        if (len < K) {
            shr_words(run, (K - len) * d, tmp);
#pragma unroll
            for (int w = 0; w < NW; ++w) run[w] &= tmp[w];
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) acc |= run[w];
    }
    return acc != 0;
}

// ------------------------------------------------------------------------------------------------
// one env held in registers
// ------------------------------------------------------------------------------------------------
template <class G>
struct EnvRegs {
    u64 pl[2][G::NW];   // [0] black, [1] white
    u32 meta;           // (move_count << 1) | current_player
};

template <class G>
MNK_DEV void env_zero(EnvRegs<G>& s) {
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int w = 0; w < G::NW; ++w) s.pl[p][w] = 0ull;
    s.meta = 0u;
}

template <class G>
MNK_DEV void env_load(const mnk_state_t& st, long long e, EnvRegs<G>& s) {
    const long long N = st.num_envs;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int w = 0; w < G::NW; ++w) s.pl[p][w] = st.bits[(size_t)(p * G::NW + w) * N + e];
    s.meta = st.meta[e];
}

template <class G>
MNK_DEV void env_store(const mnk_state_t& st, long long e, const EnvRegs<G>& s) {
    const long long N = st.num_envs;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int w = 0; w < G::NW; ++w) st.bits[(size_t)(p * G::NW + w) * N + e] = s.pl[p][w];
    st.meta[e] = s.meta;
}

struct MoveResult {
    float reward;   // 1.0 on a win of the mover, else 0.0
    bool done;      // win | draw
    bool illegal;   // cell occupied or action out of range (informational; the move is applied anyway)
};

// torch_vector_mnk_env.py:64-82, one env
template <class G>
MNK_DEV MoveResult apply_move(const G& g, EnvRegs<G>& s, long long action) {
    const int cells = g.cells();
    const u32 player = s.meta & 1u;
    const u32 count = (s.meta >> 1) + 1u;                       // :69
    const bool in_range = action >= 0 && action < cells;
    const int a = in_range ? (int)action : 0;
    const int bit = a + a / g.n();                              // row*(n+1) + col  (:64-65)
    const int wi = bit >> 6;
    const u64 one = 1ull << (bit & 63);
    bool occupied = false;
    u64 mine[G::NW];
#pragma unroll
    for (int w = 0; w < G::NW; ++w) {
        if (w == wi) occupied = ((s.pl[0][w] | s.pl[1][w]) & one) != 0;
        mine[w] = player ? s.pl[1][w] : s.pl[0][w];
        if (in_range && w == wi) mine[w] |= one;                // :68, unconditional
        if (player) s.pl[1][w] = mine[w]; else s.pl[0][w] = mine[w];
    }
    const bool win = has_line(g, mine);                         // :71
    const bool draw = (count >= (u32)cells) && !win;            // :72
    s.meta = (count << 1) | (player ^ 1u);                      // :82
    MoveResult r;
    r.reward = win ? 1.0f : 0.0f;                               // :75-77
    r.done = win || draw;                                       // :73
    r.illegal = occupied || !in_range;
    return r;
}

// only the stone placement of apply_move (no counters, no line test): what the observation / mask of the
// post-move position needs.  Lets a second warp build the views while the first one runs the rules.
template <class G>
MNK_DEV void place_stone(const G& g, EnvRegs<G>& s, long long action) {
    const bool in_range = action >= 0 && action < g.cells();
    const int a = in_range ? (int)action : 0;
    const int bit = a + a / g.n();
    const int wi = bit >> 6;
    const u64 one = in_range ? (1ull << (bit & 63)) : 0ull;
    const bool white = (s.meta & 1u) != 0u;
#pragma unroll
    for (int w = 0; w < G::NW; ++w) {
        if (w == wi) {
            if (white) s.pl[1][w] |= one; else s.pl[0][w] |= one;
        }
    }
}

// guard-strided bitboard of the board's real cells (all rows' n low bits)
template <class G>
MNK_DEV void board_mask(const G& g, u64 (&out)[G::NW]) {
#pragma unroll
    for (int w = 0; w < G::NW; ++w) out[w] = 0ull;
    const u32 row = g.n() >= 32 ? 0xffffffffu : ((1u << g.n()) - 1u);
    if constexpr (G::kStatic) {
#pragma unroll
        for (int r = 0; r < g.m(); ++r) or_field(out, r * g.stride(), row);
    } else {
        for (int r = 0; r < g.m(); ++r) or_field(out, r * g.stride(), row);
    }
}

// Dense views of one env:
//   obsd: 2*cells bits, plane `first` then the other (the reference's [2][m][n] order; swap => the
//         wrapper's canonical flip), bit index = channel*cells + r*n + c
//   legd: cells bits, 1 = empty in both planes; fix => an all-zero set gets bit 0 (wrapper :108-110)
template <class G>
MNK_DEV void build_views(const G& g, const EnvRegs<G>& s, bool swap, bool fix, u64 (&obsd)[G::NWD],
                         u64 (&legd)[G::NWL]) {
#pragma unroll
    for (int w = 0; w < G::NWD; ++w) obsd[w] = 0ull;
#pragma unroll
    for (int w = 0; w < G::NWL; ++w) legd[w] = 0ull;
    const int n = g.n(), W = g.stride(), cells = g.cells();
    const u32 rowmask = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
    auto body = [&](int r) {
        const u32 fb = get_field(s.pl[0], r * W, n);
        const u32 fw = get_field(s.pl[1], r * W, n);
        or_field(obsd, r * n, swap ? fw : fb);            // offsets stay compile-time constants
        or_field(obsd, cells + r * n, swap ? fb : fw);
        or_field(legd, r * n, ~(fb | fw) & rowmask);
    };
    if constexpr (G::kStatic) {
#pragma unroll
        for (int r = 0; r < g.m(); ++r) body(r);
    } else {
        for (int r = 0; r < g.m(); ++r) body(r);
    }
    if (fix && !any_bit(legd)) legd[0] |= 1ull;
}

// the two halves of build_views as separate functions (for kernels that build them in different warps)
template <class G>
MNK_DEV void build_obs_view(const G& g, const EnvRegs<G>& s, bool swap, u64 (&obsd)[G::NWD]) {
#pragma unroll
    for (int w = 0; w < G::NWD; ++w) obsd[w] = 0ull;
    const int n = g.n(), W = g.stride(), cells = g.cells();
    auto body = [&](int r) {
        const u32 fb = get_field(s.pl[0], r * W, n);
        const u32 fw = get_field(s.pl[1], r * W, n);
        or_field(obsd, r * n, swap ? fw : fb);
        or_field(obsd, cells + r * n, swap ? fb : fw);
    };
    if constexpr (G::kStatic) {
#pragma unroll
        for (int r = 0; r < g.m(); ++r) body(r);
    } else {
        for (int r = 0; r < g.m(); ++r) body(r);
    }
}

template <class G>
MNK_DEV void build_legal_view(const G& g, const EnvRegs<G>& s, bool fix, u64 (&legd)[G::NWL]) {
#pragma unroll
    for (int w = 0; w < G::NWL; ++w) legd[w] = 0ull;
    const int n = g.n(), W = g.stride();
    const u32 rowmask = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
    u64 occ[G::NW];
#pragma unroll
    for (int w = 0; w < G::NW; ++w) occ[w] = s.pl[0][w] | s.pl[1][w];
    auto body = [&](int r) { or_field(legd, r * n, ~get_field(occ, r * W, n) & rowmask); };
    if constexpr (G::kStatic) {
#pragma unroll
        for (int r = 0; r < g.m(); ++r) body(r);
    } else {
        for (int r = 0; r < g.m(); ++r) body(r);
    }
    if (fix && !any_bit(legd)) legd[0] |= 1ull;
}

// inverse of the obsd half of build_views (no swap): dense 2*cells bits -> guard-strided planes
template <class G>
MNK_DEV void planes_from_dense(const G& g, const u64 (&obsd)[G::NWD], EnvRegs<G>& s) {
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int w = 0; w < G::NW; ++w) s.pl[p][w] = 0ull;
    const int n = g.n(), W = g.stride(), cells = g.cells();
    auto body = [&](int r) {
        or_field(s.pl[0], r * W, get_field(obsd, r * n, n));
        or_field(s.pl[1], r * W, get_field(obsd, cells + r * n, n));
    };
    if constexpr (G::kStatic) {
#pragma unroll
        for (int r = 0; r < g.m(); ++r) body(r);
    } else {
        for (int r = 0; r < g.m(); ++r) body(r);
    }
}

// Warp-cooperative materialisation of a tile of <= 32 consecutive envs (lane L holds env e0+L's
// dense views).  Env by env, the owning lane's words are broadcast and every lane writes the
// floats / bytes at its own offset, so each store instruction covers one contiguous 256-byte
// (obs, float2 per lane) or 32-byte (mask, one byte per lane) run of the reference's
// f32[N][2][m][n] / bool[N][m*n] tensors.  Streaming stores: the tensors are write-once.
template <class G>
MNK_DEV void emit_tile(const G& g, long long e0, int tile_envs, int lane, const u64 (&obsd)[G::NWD],
                       const u64 (&legd)[G::NWL], float* __restrict__ obs, u8* __restrict__ mask) {
    const int cells = g.cells();
    const int two_cells = 2 * cells;
    for (int t = 0; t < tile_envs; ++t) {
        if (obs != nullptr) {
            float* base = obs + (size_t)(e0 + t) * two_cells;
#pragma unroll
            for (int j = 0; j < G::NWD; ++j) {
                const u64 v = __shfl_sync(MNK_FULL_WARP, obsd[j], t);
                const int pos = 64 * j + 2 * lane;
                if (pos < two_cells) {
                    const u32 two = (u32)(v >> (2 * lane)) & 3u;
                    float2 f;
                    f.x = (two & 1u) ? 1.0f : 0.0f;
                    f.y = (two & 2u) ? 1.0f : 0.0f;
                    __stcs(reinterpret_cast<float2*>(base + pos), f);
                }
            }
        }
        if (mask != nullptr) {
            u8* base = mask + (size_t)(e0 + t) * cells;
#pragma unroll
            for (int j = 0; j < 2 * G::NWL; ++j) {
                const u32 half = (j & 1) ? (u32)(legd[j >> 1] >> 32) : (u32)legd[j >> 1];
                const u32 v = __shfl_sync(MNK_FULL_WARP, half, t);
                const int pos = 32 * j + lane;
                if (pos < cells) __stcs(base + pos, (u8)((v >> lane) & 1u));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stream materialisation (static geometries, full 32-env tiles, 16-byte aligned outputs).
//
// A tile of 32 envs is 32*OB observation bits (OB = 2*cells) = exactly OB 32-bit words, and
// 32*cells mask bits = exactly `cells` words.  The CTA's compute warp (lane L owns env e0+L) stages
// every env's dense vectors in shared memory (one odd-strided row per lane); then ALL kStreamThreads
// threads of the CTA gather the rows into the two contiguous tile bitstreams and expand them with
// 128-bit stores: one nibble -> one float4 of the f32 observation (512 contiguous bytes per warp
// store), one half-word -> 16 mask bytes.  Spreading the expansion -- 84% of the kernel's
// instructions -- over 4 warps per tile keeps ~55 warps per SM resident at cfg2 (one warp per tile
// left the SM schedulers idle 2/3 of the time: ncu, profiles/).  emit_tile (warp shuffles, narrow
// stores) remains the path for tail tiles, unaligned outputs and runtime geometries.
// ------------------------------------------------------------------------------------------------
constexpr int kStreamThreads = 128;

template <class G>
struct TileStream {   // runtime geometries: unused
    static constexpr int kWords = 1;
};

template <int M, int N, int K>
struct TileStream<SGeom<M, N, K>> {
    using G = SGeom<M, N, K>;
    static constexpr int OB = 2 * M * N;             // observation bits per env == obs stream words per tile
    static constexpr int LB = M * N;                 // mask bits per env == mask stream words per tile
    static constexpr int OW = 2 * G::NWD;            // u32 chunks of obsd
    static constexpr int LW = 2 * G::NWL;            // u32 chunks of legd
    static constexpr int OSTRIDE = OW + 1;           // odd => conflict-free lane-major rows; +1 zero pad word
    static constexpr int LSTRIDE = LW + 1;
    static constexpr int kStageO = 0;
    static constexpr int kStageL = kStageO + 32 * OSTRIDE;
    static constexpr int kStreamO = kStageL + 32 * LSTRIDE;
    static constexpr int kStreamL = kStreamO + OB;
    static constexpr int kWords = kStreamL + LB;
};

// one 32-bit word of the contiguous tile bitstream, gathered from BITS-bit rows staged at `stage`
// (row stride STRIDE words, a zero pad word after each row)
template <int BITS, int STRIDE>
MNK_DEV u32 gather_stream_word(const u32* stage, int w) {
    const int bitpos = 32 * w;
    const int t = bitpos / BITS;
    const int o = bitpos - t * BITS;
    const u32* row = stage + t * STRIDE;
    u32 v;
    if constexpr (BITS >= 32) {
        v = __funnelshift_r(row[o >> 5], row[(o >> 5) + 1], o & 31);
        const int rem = BITS - o;
        if (rem < 32) v = (v & ((1u << rem) - 1u)) | (row[STRIDE] << rem);   // spill into env t+1
    } else {   // tiny boards: a 32-bit window spans several single-word rows
        v = 0u;
        int filled = 0, tt = t, oo = o;
        while (filled < 32 && tt < 32) {
            const int take = min(BITS - oo, 32 - filled);
            v |= ((stage[tt * STRIDE] >> oo) & ((1u << take) - 1u)) << filled;
            filled += take;
            ++tt;
            oo = 0;
        }
    }
    return v;
}

// Called by all kStreamThreads threads of the CTA; obsd / legd are read from warp 0 only.
template <int M, int N, int K>
MNK_DEV void emit_block_stream(const SGeom<M, N, K>&, u32* smem, long long e0, const u64 (&obsd)[SGeom<M, N, K>::NWD],
                               const u64 (&legd)[SGeom<M, N, K>::NWL], float* __restrict__ obs,
                               u8* __restrict__ mask, int stage_warp = 0, int stage_warp_mask = -1) {
    using TS = TileStream<SGeom<M, N, K>>;
    const int tid = threadIdx.x;
    if (stage_warp_mask < 0) stage_warp_mask = stage_warp;
    // 1. the warp(s) holding the views (lane L = env e0 + L) stage their 32 envs
    {
        const int srow = tid & 31;
        if (obs != nullptr && (tid >> 5) == stage_warp) {
            u32* row = smem + TS::kStageO + srow * TS::OSTRIDE;
#pragma unroll
            for (int j = 0; j < TS::OW; ++j) row[j] = (j & 1) ? (u32)(obsd[j >> 1] >> 32) : (u32)obsd[j >> 1];
            row[TS::OW] = 0u;
        }
        if (mask != nullptr && (tid >> 5) == stage_warp_mask) {
            u32* row = smem + TS::kStageL + srow * TS::LSTRIDE;
#pragma unroll
            for (int j = 0; j < TS::LW; ++j) row[j] = (j & 1) ? (u32)(legd[j >> 1] >> 32) : (u32)legd[j >> 1];
            row[TS::LW] = 0u;
        }
    }
    __syncthreads();
    // 2. tile bitstreams
    if (obs != nullptr)
        for (int w = tid; w < TS::OB; w += kStreamThreads)
            smem[TS::kStreamO + w] = gather_stream_word<TS::OB, TS::OSTRIDE>(smem + TS::kStageO, w);
    if (mask != nullptr)
        for (int w = tid; w < TS::LB; w += kStreamThreads)
            smem[TS::kStreamL + w] = gather_stream_word<TS::LB, TS::LSTRIDE>(smem + TS::kStageL, w);
    __syncthreads();
    // 3. expand
    if (obs != nullptr) {
        float4* dst = reinterpret_cast<float4*>(obs + (size_t)e0 * TS::OB);
        const u32* stream = smem + TS::kStreamO;
        const int shift = 4 * (tid & 7);
        constexpr int Q = 8 * TS::OB;                 // float4 per tile
#pragma unroll 4
        for (int q = tid; q < Q; q += kStreamThreads) {
            const u32 nib = stream[q >> 3] >> shift;
            float4 f;
            f.x = __uint_as_float((nib & 1u) * 0x3f800000u);
            f.y = __uint_as_float((nib & 2u) * 0x1fc00000u);
            f.z = __uint_as_float((nib & 4u) * 0x0fe00000u);
            f.w = __uint_as_float((nib & 8u) * 0x07f00000u);
            __stcs(dst + q, f);
        }
    }
    if (mask != nullptr) {
        uint4* dst = reinterpret_cast<uint4*>(mask + (size_t)e0 * TS::LB);
        const u32* stream = smem + TS::kStreamL;
        const int shift = 16 * (tid & 1);
        constexpr int Q = 2 * TS::LB;                 // 16-byte groups per tile
        for (int q = tid; q < Q; q += kStreamThreads) {
            const u32 h = stream[q >> 1] >> shift;
            uint4 b;
            b.x = ((h & 0xFu) * 0x00204081u) & 0x01010101u;
            b.y = (((h >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
            b.z = (((h >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
            b.w = (((h >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
            __stcs(dst + q, b);
        }
    }
}

// Tile kernels run one CTA per 32-env tile: kStreamThreads threads for static geometries (warp 0
// computes, all warps materialise), a single warp for runtime geometries.
template <class G>
constexpr int tile_cta_threads() { return G::kStatic ? kStreamThreads : 32; }

// block-uniform: can this tile take the stream path?
template <class G>
MNK_DEV bool tile_streams(int tile_envs, const float* obs, const u8* mask) {
    if constexpr (G::kStatic)
        return tile_envs == 32 && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(mask)) & 15u) == 0;
    else
        return false;
}

template <class G>
MNK_DEV void emit_block_stream_any(const G& g, u32* smem, long long e0, const u64 (&obsd)[G::NWD],
                                   const u64 (&legd)[G::NWL], float* __restrict__ obs, u8* __restrict__ mask,
                                   int stage_warp = 0, int stage_warp_mask = -1) {
    if constexpr (G::kStatic) emit_block_stream(g, smem, e0, obsd, legd, obs, mask, stage_warp, stage_warp_mask);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG; contract mirrored by oracle/mnk_oracle.py::philox4x32)
// ------------------------------------------------------------------------------------------------
#define MNK_STREAM_ACTION 0u
#define MNK_STREAM_SIDE 1u
#define MNK_STREAM_SAMPLE 2u
#define MNK_STREAM_OPPONENT 3u

MNK_DEV uint4 philox4x32_10(uint4 c, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u32 hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// counter is 64-bit: its low word is a Philox counter word, its high word is folded into the key, so a
// device-resident base (CUDA-graph replays, see mnk_masked_sample) can advance it without ever repeating
MNK_DEV uint4 mnk_philox(u64 seed, u64 global_env, u64 counter, u32 stream) {
    return philox4x32_10(make_uint4((u32)global_env, (u32)(global_env >> 32), (u32)counter, stream), (u32)seed,
                         (u32)(seed >> 32) ^ (u32)(counter >> 32));
}

// uniformly random empty cell from the bitboards (RandomPolicy.act, policy.py:17-29)
template <class G>
MNK_DEV int pick_legal(const G& g, const EnvRegs<G>& s, u32 rnd, bool deterministic) {
    u64 legal[G::NW];
    board_mask(g, legal);
    int cnt = 0;
#pragma unroll
    for (int w = 0; w < G::NW; ++w) {
        legal[w] &= ~(s.pl[0][w] | s.pl[1][w]);
        cnt += __popcll(legal[w]);
    }
    if (cnt == 0) return deterministic ? 0 : (int)__umulhi(rnd, (u32)g.cells());
    int j = deterministic ? 0 : (int)__umulhi(rnd, (u32)cnt);
    int pos = 0;
    bool found = false;
#pragma unroll
    for (int w = 0; w < G::NW; ++w) {
        const int c = __popcll(legal[w]);
        if (!found) {
            if (j < c) { pos = 64 * w + nth_set_bit64(legal[w], j); found = true; }
            else j -= c;
        }
    }
    return pos - pos / g.stride();
}
