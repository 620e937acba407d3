// mnk_env.cu -- environment kernels of libmnk_b200.so and their C-ABI entry points
// (include/mnk_b200.h).  Replaces src/env/torch_vector_mnk_env.py of the reference and
// RandomPolicy.act of src/selfplay/policy.py.
//
// Kernel shapes
//   tile kernels (step_dense, observe, pack): one warp per 32 consecutive envs.  Lane L owns env
//     e0+L for the state update (coalesced 8-byte word loads from the SoA planes), then the warp
//     materialises the 32 envs' f32 observation / bool mask rows cooperatively (emit_tile) so
//     that the dominant traffic -- 8*A + A bytes per env of output -- leaves as full contiguous runs.
//   flat kernels (reset, step_subset, meta import/export, random_legal): one thread per env.
#include "mnk_dispatch.cuh"

// strict-mode accounting: [0] += 1, [1] = max(0x7fffffff - env) i.e. encodes the SMALLEST offending env
MNK_DEV void note_illegal(int32_t* illegal, long long e) {
    atomicAdd(&illegal[0], 1);
    atomicMax(&illegal[1], (int32_t)(0x7fffffffLL - min(e, 0x7ffffffeLL)));
}

// ------------------------------------------------------------------------------------------------
// reset
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFlatThreads)
reset_kernel(mnk_state_t st, const int64_t* __restrict__ idx, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long e = idx ? idx[i] : i;
    if (e < 0 || e >= st.num_envs) return;
    const int planes = 2 * st.words;
    for (int w = 0; w < planes; ++w) st.bits[(size_t)w * st.num_envs + e] = 0ull;
    st.meta[e] = 0u;
}

// ------------------------------------------------------------------------------------------------
// dense step (+ optional observation / mask materialisation, auto-reset, strict accounting)
// ------------------------------------------------------------------------------------------------
// one dense step of this CTA's tile of 32 envs (shared by the one-step and the slab kernel)
template <class G, bool ACT32>
MNK_DEV void step_dense_tile(const G& g, const mnk_state_t& st, const void* __restrict__ actions, float* __restrict__ rewards,
                             u8* __restrict__ dones, float* __restrict__ obs, u8* __restrict__ mask,
                             int32_t* __restrict__ illegal, u32 flags, u32* tile_smem) {
    const int lane = threadIdx.x & 31;
    const long long e0 = (long long)blockIdx.x * kTileEnvs;
    const int tile_envs = (int)min((long long)kTileEnvs, st.num_envs - e0);
    const bool emit = obs != nullptr || mask != nullptr;
    const bool stream = emit && tile_streams<G>(tile_envs, obs, mask);   // block-uniform
    u64 obsd[G::NWD];
    u64 legd[G::NWL];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {   // the rules warp: lane L owns env e0 + L
        const long long e = e0 + lane;
        EnvRegs<G> s;
        env_zero(s);
        if (e < st.num_envs) {
            env_load(st, e, s);
            const long long a = ACT32 ? (long long)static_cast<const int32_t*>(actions)[e]
                                      : (long long)static_cast<const int64_t*>(actions)[e];
            const MoveResult r = apply_move(g, s, a);
            rewards[e] = r.reward;
            dones[e] = r.done ? 1 : 0;
            if (illegal != nullptr && r.illegal) note_illegal(illegal, e);
            if (emit && !stream) build_views(g, s, false, false, obsd, legd);
            if ((flags & MNK_STEP_AUTORESET) && r.done) env_zero(s);
            if (stream) asm volatile("bar.sync 2, 96;" ::: "memory");   // the views warps have read the pre-move state
            env_store(st, e, s);
        }
        if (emit && !stream) emit_tile(g, e0, tile_envs, lane, obsd, legd, obs, mask);
    } else if (warp <= 2 && stream) {
        // the views warps (stream path only: full 32-env tile): in parallel with the rules warp they read the
        // tile's PRE-move state, place the stone and build the dense views -- warp 1 the observation bits,
        // warp 2 the legal-cell bits -- which takes ~240 of the ~400 serial instructions off the tile's critical
        // path.  Named barrier 2 orders their loads before warp 0's store.
        const long long e = e0 + lane;
        EnvRegs<G> s;
        env_load(st, e, s);
        const long long a = ACT32 ? (long long)static_cast<const int32_t*>(actions)[e]
                                  : (long long)static_cast<const int64_t*>(actions)[e];
        asm volatile("bar.arrive 2, 96;" ::: "memory");
        place_stone(g, s, a);
        if (warp == 1) build_obs_view(g, s, false, obsd);
        else build_legal_view(g, s, false, legd);
    }
    if (stream) emit_block_stream_any(g, tile_smem, e0, obsd, legd, obs, mask, 1, 2);
}

template <class G, bool ACT32>
__global__ void __launch_bounds__(tile_cta_threads<G>())
step_dense_kernel(G g, mnk_state_t st, const void* __restrict__ actions, float* __restrict__ rewards,
                  u8* __restrict__ dones, float* __restrict__ obs, u8* __restrict__ mask,
                  int32_t* __restrict__ illegal, u32 flags) {
    __shared__ u32 tile_smem[TileStream<G>::kWords];
    // Programmatic dependent launch (MNK_STEP_PDL): let the next step's CTAs become resident while this grid
    // drains, and do not touch anything the previous step wrote before it has completed.  Both are no-ops
    // for a plain launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    step_dense_tile<G, ACT32>(g, st, actions, rewards, dones, obs, mask, illegal, flags, tile_smem);
}

// A SLAB of consecutive dense steps in ONE launch (mnk_step_slab): envs are independent and a CTA owns its tile of 32
// for the whole launch, so step s + 1 of a tile only needs step s of the same tile -- a block barrier, not a grid-wide
// one.  Same device work as `steps` launches of step_dense_kernel (each step materialises its own observation / mask),
// one launch's worth of host work and no fill / drain between the steps.
constexpr int kMaxSlabSteps = MNK_MAX_SLAB_STEPS;
struct SlabArgs {
    int steps;
    long long action_stride, rd_stride;      // bytes between consecutive steps' action batches / reward + done blocks
    float* obs[kMaxSlabSteps];
    u8* mask[kMaxSlabSteps];
};

template <class G, bool ACT32>
__global__ void __launch_bounds__(tile_cta_threads<G>())
step_dense_slab_kernel(G g, mnk_state_t st, const char* __restrict__ actions, char* __restrict__ rd, SlabArgs sa, u32 flags) {
    __shared__ u32 tile_smem[TileStream<G>::kWords];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int s = 0; s < sa.steps; ++s) {
        char* block = rd + (size_t)s * sa.rd_stride;
        step_dense_tile<G, ACT32>(g, st, actions + (size_t)s * sa.action_stride, reinterpret_cast<float*>(block),
                                  reinterpret_cast<u8*>(block) + 4 * (size_t)st.num_envs, sa.obs[s], sa.mask[s], nullptr, flags,
                                  tile_smem);
        __syncthreads();      // the tile's state written by this step is visible to every warp of the CTA; staging reusable
    }
}

// ------------------------------------------------------------------------------------------------
// step_subset: scattered envs, no materialisation (mnk_step runs observe afterwards)
// ------------------------------------------------------------------------------------------------
template <class G, bool ACT32>
__global__ void __launch_bounds__(kFlatThreads)
step_subset_kernel(G g, mnk_state_t st, const void* __restrict__ actions, const int64_t* __restrict__ idx,
                   long long n_active, float* __restrict__ rewards, u8* __restrict__ dones,
                   int32_t* __restrict__ illegal, u32 flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    const long long e = idx[i];
    if (e < 0 || e >= st.num_envs) return;
    EnvRegs<G> s;
    env_load(st, e, s);
    const long long a = ACT32 ? (long long)static_cast<const int32_t*>(actions)[i]
                              : (long long)static_cast<const int64_t*>(actions)[i];
    const MoveResult r = apply_move(g, s, a);
    rewards[e] = r.reward;
    dones[e] = r.done ? 1 : 0;
    if (illegal != nullptr && r.illegal) note_illegal(illegal, e);
    if ((flags & MNK_STEP_AUTORESET) && r.done) env_zero(s);
    env_store(st, e, s);
}

// ------------------------------------------------------------------------------------------------
// observe / unpack
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void __launch_bounds__(tile_cta_threads<G>())
observe_kernel(G g, mnk_state_t st, float* __restrict__ obs, u8* __restrict__ mask,
               const u8* __restrict__ swap, int fix_all_masked) {
    __shared__ u32 tile_smem[TileStream<G>::kWords];
    const int lane = threadIdx.x & 31;
    const long long e0 = (long long)blockIdx.x * kTileEnvs;
    const int tile_envs = (int)min((long long)kTileEnvs, st.num_envs - e0);
    const bool stream = tile_streams<G>(tile_envs, obs, mask);   // block-uniform
    u64 obsd[G::NWD];
    u64 legd[G::NWL];
    const int warp = threadIdx.x >> 5;
    if (stream) {
        // full tile on the stream path: warp 0 builds the observation bits, warp 1 the legal-cell bits, concurrently
        // (the same split as step_dense_kernel's view warps); both read the state, nobody writes it
        if ((warp == 0 && obs != nullptr) || (warp == 1 && mask != nullptr)) {
            const long long e = e0 + lane;
            EnvRegs<G> s;
            env_load(st, e, s);
            if (warp == 0) build_obs_view(g, s, swap != nullptr && swap[e] != 0, obsd);
            else build_legal_view(g, s, fix_all_masked != 0, legd);
        }
        emit_block_stream_any(g, tile_smem, e0, obsd, legd, obs, mask, 0, 1);
    } else if (warp == 0) {
        const long long e = e0 + lane;
        EnvRegs<G> s;
        env_zero(s);
        bool sw = false;
        if (e < st.num_envs) {
            env_load(st, e, s);
            sw = swap != nullptr && swap[e] != 0;
        }
        build_views(g, s, sw, fix_all_masked != 0, obsd, legd);
        emit_tile(g, e0, tile_envs, lane, obsd, legd, obs, mask);
    }
}

// ------------------------------------------------------------------------------------------------
// pack: f32[N][2][m][n] (non-zero = stone) -> bitboards; coalesced reads + ballot transpose
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void __launch_bounds__(kTileThreads)
pack_kernel(G g, mnk_state_t st, const float* __restrict__ boards) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * kTileWarps + (threadIdx.x >> 5);
    const long long e0 = tile * kTileEnvs;
    if (e0 >= st.num_envs) return;
    const int tile_envs = (int)min((long long)kTileEnvs, st.num_envs - e0);
    const int two_cells = 2 * g.cells();
    u64 obsd[G::NWD];
#pragma unroll
    for (int w = 0; w < G::NWD; ++w) obsd[w] = 0ull;
    for (int t = 0; t < tile_envs; ++t) {
        const float* base = boards + (size_t)(e0 + t) * two_cells;
#pragma unroll
        for (int j = 0; j < 2 * G::NWD; ++j) {
            const int pos = 32 * j + lane;
            const float v = pos < two_cells ? __ldg(base + pos) : 0.0f;
            const u32 b = __ballot_sync(MNK_FULL_WARP, v != 0.0f);
            if (lane == t) obsd[j >> 1] |= (u64)b << ((j & 1) * 32);
        }
    }
    const long long e = e0 + lane;
    if (e < st.num_envs) {
        EnvRegs<G> s;
        planes_from_dense(g, obsd, s);
        const long long N = st.num_envs;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int w = 0; w < G::NW; ++w) st.bits[(size_t)(p * G::NW + w) * N + e] = s.pl[p][w];
    }
}

// ------------------------------------------------------------------------------------------------
// meta import / export
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFlatThreads)
export_meta_kernel(mnk_state_t st, int64_t* __restrict__ player, int64_t* __restrict__ count) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.num_envs) return;
    const u32 meta = st.meta[e];
    if (player) player[e] = meta & 1u;
    if (count) count[e] = meta >> 1;
}

__global__ void __launch_bounds__(kFlatThreads)
import_meta_kernel(mnk_state_t st, const int64_t* __restrict__ player, const int64_t* __restrict__ count) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.num_envs) return;
    u32 meta = st.meta[e];
    if (player) meta = (meta & ~1u) | (u32)(player[e] & 1);
    if (count) meta = (meta & 1u) | ((u32)count[e] << 1);
    st.meta[e] = meta;
}

// ------------------------------------------------------------------------------------------------
// RandomPolicy.act from the bitboards
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void __launch_bounds__(kFlatThreads)
random_legal_kernel(G g, mnk_state_t st, u64 seed, u64 counter, long long env_offset, int deterministic,
                    int64_t* __restrict__ actions) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.num_envs) return;
    EnvRegs<G> s;
    env_load(st, e, s);
    const u32 rnd = deterministic ? 0u : mnk_philox(seed, (u64)(env_offset + e), counter, MNK_STREAM_ACTION).x;
    actions[e] = pick_legal(g, s, rnd, deterministic != 0);
}

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int mnk_version(void) { return MNK_B200_VERSION; }

const char* mnk_error_string(int code) {
    switch (code) {
        case MNK_OK: return "ok";
        case MNK_ERR_NULL: return "a required pointer is NULL";
        case MNK_ERR_GEOM: return "unsupported board geometry (need 1 <= k <= min(m, n), n <= 32, m*(n+1) <= 512)";
        case MNK_ERR_ALIGN: return "buffer misaligned";
        case MNK_ERR_ARG: return "invalid argument";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int mnk_state_words(int m, int n) { return mnk_words_for(m, n); }

int mnk_reset(const mnk_state_t* st, const int64_t* idx, int64_t n_idx, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (idx == nullptr) {
        if (st->num_envs == 0) return MNK_OK;
        cudaError_t e = cudaMemsetAsync(st->bits, 0, sizeof(u64) * 2 * st->words * (size_t)st->num_envs, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(st->meta, 0, sizeof(u32) * (size_t)st->num_envs, s);
        return e == cudaSuccess ? MNK_OK : (int)e;
    }
    if (n_idx < 0) return MNK_ERR_ARG;
    if (n_idx == 0) return MNK_OK;
    reset_kernel<<<mnk_flat_blocks(n_idx), kFlatThreads, 0, s>>>(*st, idx, n_idx);
    return mnk_launch_status();
}

int mnk_observe(const mnk_state_t* st, float* obs, uint8_t* mask, const uint8_t* swap, int fix_all_masked,
                void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (obs == nullptr && mask == nullptr) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(obs) & 7u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(*st, [&](auto g) {
        observe_kernel<<<mnk_cta_tiles(st->num_envs), tile_cta_threads<decltype(g)>(), 0, s>>>(g, *st, obs, mask, swap, fix_all_masked);
        return mnk_launch_status();
    });
}

int mnk_unpack_boards(const mnk_state_t* st, float* boards, void* stream) {
    if (boards == nullptr) return MNK_ERR_NULL;
    return mnk_observe(st, boards, nullptr, nullptr, 0, stream);
}

int mnk_pack_boards(const mnk_state_t* st, const float* boards, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (boards == nullptr) return MNK_ERR_NULL;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(*st, [&](auto g) {
        pack_kernel<<<mnk_tile_blocks(st->num_envs), kTileThreads, 0, s>>>(g, *st, boards);
        return mnk_launch_status();
    });
}

int mnk_step_slab(const mnk_state_t* st, const void* actions, int64_t action_stride, void* rewards_dones, int64_t rd_stride,
                  int32_t steps, float* const* obs, uint8_t* const* mask, uint32_t flags, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (actions == nullptr || rewards_dones == nullptr) return MNK_ERR_NULL;
    if (steps < 0 || steps > kMaxSlabSteps || rd_stride < 5 * st->num_envs || (rd_stride & 3) != 0) return MNK_ERR_ARG;
    if (action_stride < (int64_t)((flags & MNK_STEP_ACTIONS_I32) ? 4 : 8) * st->num_envs) return MNK_ERR_ARG;
    if (steps == 0 || st->num_envs == 0) return MNK_OK;
    SlabArgs sa;
    sa.steps = steps; sa.action_stride = action_stride; sa.rd_stride = rd_stride;
    bool views = false;
    for (int i = 0; i < kMaxSlabSteps; ++i) {
        sa.obs[i] = (obs != nullptr && i < steps) ? obs[i] : nullptr;
        sa.mask[i] = (mask != nullptr && i < steps) ? mask[i] : nullptr;
        if (reinterpret_cast<uintptr_t>(sa.obs[i]) & 7u) return MNK_ERR_ALIGN;
        views = views || sa.obs[i] != nullptr || sa.mask[i] != nullptr;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool act32 = (flags & MNK_STEP_ACTIONS_I32) != 0;
    return mnk_dispatch_geom(*st, [&](auto g) {
        using G = decltype(g);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(mnk_cta_tiles(st->num_envs));
        cfg.blockDim = dim3(views ? tile_cta_threads<G>() : 32);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (flags & MNK_STEP_PDL) ? 1 : 0;
        const mnk_state_t stv = *st;
        const char* a = static_cast<const char*>(actions);
        char* rd = static_cast<char*>(rewards_dones);
        cudaError_t e;
        if (act32) e = cudaLaunchKernelEx(&cfg, step_dense_slab_kernel<G, true>, g, stv, a, rd, sa, flags);
        else e = cudaLaunchKernelEx(&cfg, step_dense_slab_kernel<G, false>, g, stv, a, rd, sa, flags);
        return e == cudaSuccess ? mnk_launch_status() : (int)e;
    });
}

int mnk_step(const mnk_state_t* st, const void* actions, const int64_t* idx, int64_t n_active, float* rewards,
             uint8_t* dones, float* obs, uint8_t* mask, int32_t* illegal, uint32_t flags, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (actions == nullptr || rewards == nullptr || dones == nullptr) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(obs) & 7u) return MNK_ERR_ALIGN;
    if (n_active < 0 || (idx == nullptr && n_active != st->num_envs)) return MNK_ERR_ARG;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool act32 = (flags & MNK_STEP_ACTIONS_I32) != 0;
    if (idx == nullptr) {
        return mnk_dispatch_geom(*st, [&](auto g) {
            using G = decltype(g);
            const unsigned blocks = mnk_cta_tiles(st->num_envs);
            // no materialisation => only the compute warp has work: launch single-warp CTAs
            const int threads = (obs != nullptr || mask != nullptr) ? tile_cta_threads<G>() : 32;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(blocks);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = (flags & MNK_STEP_PDL) ? 1 : 0;
            const mnk_state_t stv = *st;
            cudaError_t e;
            if (act32)
                e = cudaLaunchKernelEx(&cfg, step_dense_kernel<G, true>, g, stv, actions, rewards, dones, obs, mask, illegal, flags);
            else
                e = cudaLaunchKernelEx(&cfg, step_dense_kernel<G, false>, g, stv, actions, rewards, dones, obs, mask, illegal, flags);
            return e == cudaSuccess ? mnk_launch_status() : (int)e;
        });
    }
    // step_subset: full-size zeroed outputs, scattered update, then observe() over all envs
    cudaError_t e = cudaMemsetAsync(rewards, 0, sizeof(float) * (size_t)st->num_envs, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dones, 0, (size_t)st->num_envs, s);
    if (e != cudaSuccess) return (int)e;
    if (n_active > 0) {
        const int rc = mnk_dispatch_geom(*st, [&](auto g) {
            using G = decltype(g);
            const unsigned blocks = mnk_flat_blocks(n_active);
            if (act32)
                step_subset_kernel<G, true><<<blocks, kFlatThreads, 0, s>>>(g, *st, actions, idx, n_active, rewards, dones, illegal, flags);
            else
                step_subset_kernel<G, false><<<blocks, kFlatThreads, 0, s>>>(g, *st, actions, idx, n_active, rewards, dones, illegal, flags);
            return mnk_launch_status();
        });
        if (rc != MNK_OK) return rc;
    }
    if (obs != nullptr || mask != nullptr) return mnk_observe(st, obs, mask, nullptr, 0, stream);
    return MNK_OK;
}

int mnk_step_host(const mnk_state_t* st, const void* host_actions, void* dev_actions, void* dev_rd, void* host_rd,
                  float* obs, uint8_t* mask, uint32_t flags, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (host_actions == nullptr || host_rd == nullptr) return MNK_ERR_NULL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)st->num_envs;
    if (flags & MNK_STEP_ZEROCOPY) {   // the kernel reads the pinned actions and writes rewards / dones over PCIe itself
        float* rewards_h = static_cast<float*>(host_rd);
        uint8_t* dones_h = static_cast<uint8_t*>(host_rd) + 4 * n;
        const int rc = mnk_step(st, host_actions, nullptr, st->num_envs, rewards_h, dones_h, obs, mask, nullptr,
                                flags & ~(MNK_STEP_ZEROCOPY | MNK_STEP_NOSYNC), stream);
        if (rc != MNK_OK || (flags & MNK_STEP_NOSYNC)) return rc;
        const cudaError_t e = cudaStreamSynchronize(s);
        return e == cudaSuccess ? MNK_OK : (int)e;
    }
    if (dev_actions == nullptr || dev_rd == nullptr) return MNK_ERR_NULL;
    const size_t abytes = n * ((flags & MNK_STEP_ACTIONS_I32) ? 4 : 8);
    cudaError_t e = cudaMemcpyAsync(dev_actions, host_actions, abytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    float* rewards = static_cast<float*>(dev_rd);
    uint8_t* dones = static_cast<uint8_t*>(dev_rd) + 4 * n;
    const int rc = mnk_step(st, dev_actions, nullptr, st->num_envs, rewards, dones, obs, mask, nullptr, flags & ~MNK_STEP_NOSYNC, stream);
    if (rc != MNK_OK) return rc;
    e = cudaMemcpyAsync(host_rd, dev_rd, 5 * n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && !(flags & MNK_STEP_NOSYNC)) e = cudaStreamSynchronize(s);
    return e == cudaSuccess ? MNK_OK : (int)e;
}

int mnk_export_meta(const mnk_state_t* st, int64_t* current_player, int64_t* move_counts, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (st->num_envs == 0) return MNK_OK;
    export_meta_kernel<<<mnk_flat_blocks(st->num_envs), kFlatThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        *st, current_player, move_counts);
    return mnk_launch_status();
}

int mnk_import_meta(const mnk_state_t* st, const int64_t* current_player, const int64_t* move_counts, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (st->num_envs == 0) return MNK_OK;
    import_meta_kernel<<<mnk_flat_blocks(st->num_envs), kFlatThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        *st, current_player, move_counts);
    return mnk_launch_status();
}

int mnk_random_legal(const mnk_state_t* st, uint64_t seed, uint64_t counter, int64_t env_offset, int deterministic,
                     int64_t* actions, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (actions == nullptr) return MNK_ERR_NULL;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(*st, [&](auto g) {
        random_legal_kernel<<<mnk_flat_blocks(st->num_envs), kFlatThreads, 0, s>>>(g, *st, seed, counter, env_offset,
                                                                                 deterministic, actions);
        return mnk_launch_status();
    });
}

}  // extern "C"
