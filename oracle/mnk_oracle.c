/*
 * mnk_oracle.c -- CPU ORACLE for the batched MNK environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the
 * reference algorithm (michal-szadkowski/rl-selfplay-mnk,
 * src/env/torch_vector_mnk_env.py).  It is the checker the CUDA path is
 * compared against.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it; the product path (rl-selfplay-mnk_b200/) never
 * does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
 * against tests/golden/*.npz, which oracle/gen_golden.py produced by importing
 * and running the unmodified reference in the build container.
 *
 * State layout deliberately mirrors the reference (one occupancy byte per cell
 * standing in for its f32 0/1 one-hot planes), NOT the bitboards of the CUDA
 * path, so that the two implementations share no representation.
 *
 *   boards          u8 [N][2][m][n]   plane 0 = black, plane 1 = white
 *   current_player  i64[N]
 *   move_counts     i64[N]
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline uint8_t* plane_of(uint8_t* boards, int m, int n, int64_t env, int64_t player) {
    return boards + ((size_t)env * 2 + (size_t)player) * (size_t)m * (size_t)n;
}

/* reference: torch_vector_mnk_env.py:34-42  (reset; the observe() it ends with
 * is a separate call here). idx == NULL <=> env_indices is None. */
ORC_API void orc_reset(uint8_t* boards, int64_t* current_player, int64_t* move_counts,
                       int m, int n, int64_t num_envs, const int64_t* idx, int64_t n_idx) {
    const size_t per_env = (size_t)2 * m * n;
    if (idx == NULL) {
        memset(boards, 0, per_env * (size_t)num_envs);
        memset(current_player, 0, sizeof(int64_t) * (size_t)num_envs);
        memset(move_counts, 0, sizeof(int64_t) * (size_t)num_envs);
        return;
    }
    for (int64_t i = 0; i < n_idx; ++i) {
        const int64_t e = idx[i];
        memset(boards + per_env * (size_t)e, 0, per_env);
        current_player[e] = 0; /* PLAYER_BLACK, constants.py:1 */
        move_counts[e] = 0;
    }
}

/* reference: torch_vector_mnk_env.py:46-53.  observation = raw copy of the planes
 * as f32, action_mask[e][a] = cell a empty in BOTH planes. */
ORC_API void orc_observe(const uint8_t* boards, int m, int n, int64_t num_envs,
                         float* observation, uint8_t* action_mask) {
    const int cells = m * n;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < num_envs; ++e) {
        const uint8_t* b = boards + (size_t)e * 2 * cells;
        if (observation) {
            float* o = observation + (size_t)e * 2 * cells;
            for (int i = 0; i < 2 * cells; ++i) o[i] = (float)b[i];
        }
        if (action_mask) {
            uint8_t* am = action_mask + (size_t)e * cells;
            for (int i = 0; i < cells; ++i) am[i] = (uint8_t)!(b[i] != 0 || b[cells + i] != 0);
        }
    }
}

/* reference: torch_vector_mnk_env.py:106-119 (+ the kernels of :26-32).
 * Three "valid" cross-correlations over the mover's plane:
 *   ones[1,k]            -> outputs (m)     x (n-k+1)
 *   ones[k,1]            -> outputs (m-k+1) x (n)
 *   eye(k), fliplr(eye)  -> outputs (m-k+1) x (n-k+1), two channels
 * an output wins when its sum exceeds k - 0.1. */
ORC_API int orc_plane_has_line(const uint8_t* plane, int m, int n, int k) {
    const float threshold = (float)k - 0.1f;
    for (int r = 0; r < m; ++r)
        for (int c = 0; c + k <= n; ++c) {
            float s = 0.f;
            for (int t = 0; t < k; ++t) s += (float)plane[r * n + c + t];
            if (s > threshold) return 1;
        }
    for (int r = 0; r + k <= m; ++r)
        for (int c = 0; c < n; ++c) {
            float s = 0.f;
            for (int t = 0; t < k; ++t) s += (float)plane[(r + t) * n + c];
            if (s > threshold) return 1;
        }
    for (int r = 0; r + k <= m; ++r)
        for (int c = 0; c + k <= n; ++c) {
            float s_main = 0.f, s_anti = 0.f;
            for (int t = 0; t < k; ++t) {
                s_main += (float)plane[(r + t) * n + c + t];         /* eye(k)          */
                s_anti += (float)plane[(r + t) * n + c + (k - 1 - t)]; /* fliplr(eye(k)) */
            }
            if (s_main > threshold || s_anti > threshold) return 1;
        }
    return 0;
}

/* reference: torch_vector_mnk_env.py:60-84 (step_subset; step() is the same call
 * with idx = arange(N), :58).  rewards f32[N] / dones u8[N] are FULL SIZE and
 * zero for unlisted envs.  No legality, bounds or terminal check (:86-104 are dead
 * code in the reference).  Actions outside [0, m*n) are undefined in the reference
 * (index error); this restatement places no stone for them. */
ORC_API void orc_step_subset(uint8_t* boards, int64_t* current_player, int64_t* move_counts,
                             int m, int n, int k, int64_t num_envs,
                             const int64_t* actions, const int64_t* idx, int64_t n_idx,
                             float* rewards, uint8_t* dones) {
    const int64_t max_moves = (int64_t)m * n;
    memset(rewards, 0, sizeof(float) * (size_t)num_envs);
    memset(dones, 0, (size_t)num_envs);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_idx; ++i) {
        const int64_t e = idx ? idx[i] : i;
        const int64_t a = actions[i];
        const int64_t player = current_player[e];              /* :67 */
        uint8_t* mine = plane_of(boards, m, n, e, player);
        if (a >= 0 && a < max_moves) {
            const int64_t row = a / n, col = a % n;            /* :64-65 */
            mine[row * n + col] = 1;                           /* :68 */
        }
        move_counts[e] += 1;                                   /* :69 */
        const int win = orc_plane_has_line(mine, m, n, k);     /* :71 */
        const int draw = (move_counts[e] >= max_moves) && !win; /* :72 */
        if (win) rewards[e] = 1.0f;                            /* :75-77 */
        dones[e] = (uint8_t)(win || draw);                     /* :73,79-80 */
        current_player[e] = player ^ 1;                        /* :82 */
    }
}

/* Helper for test drivers: number of empty cells and the j-th empty cell (ascending
 * cell index) of one env; -1 when j is out of range.  Not part of the reference. */
ORC_API int64_t orc_nth_legal(const uint8_t* boards, int m, int n, int64_t env, int64_t j) {
    const int cells = m * n;
    const uint8_t* b = boards + (size_t)env * 2 * cells;
    int64_t seen = 0;
    for (int i = 0; i < cells; ++i)
        if (!(b[i] || b[cells + i])) {
            if (seen == j) return i;
            ++seen;
        }
    return -1;
}

ORC_API int64_t orc_count_legal(const uint8_t* boards, int m, int n, int64_t env) {
    const int cells = m * n;
    const uint8_t* b = boards + (size_t)env * 2 * cells;
    int64_t cnt = 0;
    for (int i = 0; i < cells; ++i) cnt += !(b[i] || b[cells + i]);
    return cnt;
}
