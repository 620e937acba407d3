// mnk_dispatch.cuh -- host-side helpers: argument validation, geometry dispatch, launch shapes.
#pragma once
#include <atomic>

#include "mnk_device.cuh"

// Geometries compiled with every loop bound and shift distance constant.  Anything else runs the
// same templates with a runtime geometry (DGeom<words>).
#define MNK_STATIC_GEOMS(X) X(3, 3, 3) X(9, 9, 5) X(13, 13, 5) X(15, 15, 5) X(19, 19, 5)

static inline int mnk_words_for(int m, int n) {
    if (m < 1 || n < 1 || n > 32) return MNK_ERR_GEOM;
    const long long bits = (long long)m * (n + 1);
    if (bits > 64LL * MNK_MAX_WORDS) return MNK_ERR_GEOM;
    return (int)((bits + 63) / 64);
}

static inline int mnk_check_state(const mnk_state_t* st) {
    if (st == nullptr || st->bits == nullptr || st->meta == nullptr) return MNK_ERR_NULL;
    const int words = mnk_words_for(st->m, st->n);
    if (words < 0) return words;
    if (st->k < 1 || st->k > st->m || st->k > st->n) return MNK_ERR_GEOM;
    if (words != st->words || st->num_envs < 0) return MNK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(st->bits) & 7u) || (reinterpret_cast<uintptr_t>(st->meta) & 3u))
        return MNK_ERR_ALIGN;
    return MNK_OK;
}

// calls f(geom) with the matching SGeom<> / DGeom<> value; f returns int
template <class F>
static inline int mnk_dispatch_geom(const mnk_state_t& st, F&& f) {
#define MNK_TRY_STATIC(M, N, K) \
    if (st.m == M && st.n == N && st.k == K) return f(SGeom<M, N, K>{});
    MNK_STATIC_GEOMS(MNK_TRY_STATIC)
#undef MNK_TRY_STATIC
    switch (st.words) {
        case 1: return f(DGeom<1>{st.m, st.n, st.k});
        case 2: return f(DGeom<2>{st.m, st.n, st.k});
        case 3: return f(DGeom<3>{st.m, st.n, st.k});
        case 4: return f(DGeom<4>{st.m, st.n, st.k});
        case 5: return f(DGeom<5>{st.m, st.n, st.k});
        case 6: return f(DGeom<6>{st.m, st.n, st.k});
        case 7: return f(DGeom<7>{st.m, st.n, st.k});
        case 8: return f(DGeom<8>{st.m, st.n, st.k});
        default: return MNK_ERR_GEOM;
    }
}

static inline int mnk_launch_status() {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MNK_OK : (int)e;
}

// Opt-in dynamic shared memory is a per-device function attribute.  `granted` (one slot per device,
// zero-initialised by the caller as a function-local static) remembers the largest size already set on
// each device, so a process driving several GPUs configures each of them and concurrent callers at
// worst repeat an idempotent call.
constexpr int kMaxDevices = 64;
template <class Kernel>
static inline int mnk_optin_smem(Kernel kernel, size_t bytes, std::atomic<size_t>* granted) {
    if (bytes <= 48 * 1024) return MNK_OK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (tracked && granted[dev].load(std::memory_order_relaxed) >= bytes) return MNK_OK;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    if (tracked) granted[dev].store(bytes, std::memory_order_relaxed);
    return MNK_OK;
}

// SM count of the current device (cached per device; 148 on B200) for persistent grids
static inline int mnk_sm_count() {
    static std::atomic<int> cached[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (tracked) {
        const int c = cached[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    if (tracked) cached[dev].store(sms, std::memory_order_relaxed);
    return sms;
}

// warp-tile kernels (pack): one warp per 32 consecutive envs, 4 warps per CTA
constexpr int kTileEnvs = 32;
constexpr int kTileWarps = 4;
constexpr int kTileThreads = kTileWarps * 32;
static inline unsigned mnk_tile_blocks(long long num_envs) {
    const long long tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
    return (unsigned)((tiles + kTileWarps - 1) / kTileWarps);
}
// CTA-per-tile kernels (step_dense, observe): one CTA per 32 consecutive envs
static inline unsigned mnk_cta_tiles(long long num_envs) { return (unsigned)((num_envs + kTileEnvs - 1) / kTileEnvs); }
// thread-per-env kernels
constexpr int kFlatThreads = 256;
static inline unsigned mnk_flat_blocks(long long count) { return (unsigned)((count + kFlatThreads - 1) / kFlatThreads); }
