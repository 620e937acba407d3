// mnk_rollout.cu -- rollout storage for PPO on packed bitboards.
//
// The reference's RolloutBuffer (src/alg/rollout_buffer.py:13-58) keeps f32 observations
// [T][N][2][m][n] and bool masks [T][N][m*n]: 729 B per agent-step at 9x9, ~102 GB for BASELINE cfg3
// (512 x 262,144).  Here a step's observation is the agent's two canonical planes as guard-strided
// bitboards (32 B at 9x9, same SoA layout as mnk_state.bits, one slot per step); the mask is
// derived on the fly.  Minibatches are materialised back to the reference's f32 / bool tensors only
// at gather time, for the rows a minibatch actually uses.
//
//   mnk_rollout_store_obs   <- RolloutBuffer.add, observation + action_mask part (:51,:57)
//   mnk_rollout_gather      <- RolloutBuffer.get_data_loader, b_obs[batch_idx] / b_masks[batch_idx] (:101-110)
//   mnk_gae                 <- RolloutBuffer.compute_advantages_and_returns (:60-80)
//   mnk_episode_stats       <- PPOAgent.learn episode accounting (src/alg/ppo.py:110-120) without .tolist()
#include "mnk_dispatch.cuh"

// canonical planes of the current state: me = the agent's stones, enemy = the opponent's
__global__ void __launch_bounds__(kFlatThreads)
store_obs_kernel(mnk_state_t st, const u8* __restrict__ agent_side, u64* __restrict__ slot) {
    const u64* bits = reinterpret_cast<const u64*>(st.bits);
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.num_envs) return;
    const long long N = st.num_envs;
    const bool white = agent_side != nullptr && agent_side[e] != 0;
    for (int w = 0; w < st.words; ++w) {
        const u64 b = bits[(size_t)w * N + e];
        const u64 wh = bits[(size_t)(st.words + w) * N + e];
        slot[(size_t)w * N + e] = white ? wh : b;
        slot[(size_t)(st.words + w) * N + e] = white ? b : wh;
    }
}

// minibatch gather: sample i of the flattened [T*N] rollout -> f32 observation + bool mask rows
template <class G>
__global__ void __launch_bounds__(tile_cta_threads<G>())
rollout_gather_kernel(G g, const u64* __restrict__ packed, long long num_envs, const int64_t* __restrict__ index,
                      long long count, float* __restrict__ obs, u8* __restrict__ mask) {
    __shared__ u32 tile_smem[TileStream<G>::kWords];
    const int lane = threadIdx.x & 31;
    const long long r0 = (long long)blockIdx.x * kTileEnvs;
    const int tile_rows = (int)min((long long)kTileEnvs, count - r0);
    const bool stream = tile_streams<G>(tile_rows, obs, mask);
    u64 obsd[G::NWD];
    u64 legd[G::NWL];
    if (threadIdx.x < 32) {
        const long long r = r0 + lane;
        EnvRegs<G> s;
        env_zero(s);
        if (r < count) {
            const long long i = index ? index[r] : r;
            const long long t = i / num_envs, e = i - t * num_envs;
            const u64* slot = packed + (size_t)t * 2 * G::NW * num_envs;
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int w = 0; w < G::NW; ++w) s.pl[p][w] = slot[(size_t)(p * G::NW + w) * num_envs + e];
        }
        build_views(g, s, false, true, obsd, legd);     // stored planes are already canonical; all-masked fix as in :108-110
        if (!stream) emit_tile(g, r0, tile_rows, lane, obsd, legd, obs, mask);
    }
    if (stream) emit_block_stream_any(g, tile_smem, r0, obsd, legd, obs, mask);
}

// GAE(lambda), one thread per env, reverse scan over t.  Same operation order as the reference's
// tensor expressions (no fused multiply-add) so results are bit-identical to it in fp32.
__global__ void __launch_bounds__(kFlatThreads)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const u8* __restrict__ dones,
           const float* __restrict__ last_values, long long steps, long long num_envs, float gamma, float gl,
           float* __restrict__ advantages, float* __restrict__ returns) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= num_envs) return;
    float last_gae = 0.0f;
    float next_value = last_values[e];
    // gl = (float)(gamma * lambda) with the product taken in DOUBLE on the host: the reference multiplies the two
    // Python floats first and only then meets an fp32 tensor (rollout_buffer.py:76)
    for (long long t = steps - 1; t >= 0; --t) {
        const size_t i = (size_t)t * num_envs + e;
        const float nnt = 1.0f - (dones[i] ? 1.0f : 0.0f);
        const float v = values[i];
        // delta = r + gamma * next_v * nnt - v ;  gae = delta + gamma * lambda * nnt * gae
        const float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, next_value), nnt)), v);
        last_gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), last_gae));
        advantages[i] = last_gae;
        returns[i] = __fadd_rn(last_gae, v);
        next_value = v;
    }
}

// per-step episode accounting on the device: totals = {episodes, sum reward, sum length, wins, losses, draws}
__global__ void __launch_bounds__(kFlatThreads)
episode_stats_kernel(const float* __restrict__ rewards, const u8* __restrict__ dones, long long num_envs,
                     float* __restrict__ ep_reward, float* __restrict__ ep_len, double* __restrict__ totals) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float fin = 0.f, rsum = 0.f, lsum = 0.f, win = 0.f, loss = 0.f, draw = 0.f;
    if (e < num_envs) {
        const float r = rewards[e];
        const float er = ep_reward[e] + r;
        const float el = ep_len[e] + 1.0f;
        if (dones[e]) {
            fin = 1.f; rsum = er; lsum = el;
            win = r > 0.5f; loss = r < -0.5f; draw = (r <= 0.5f && r >= -0.5f);
            ep_reward[e] = 0.f; ep_len[e] = 0.f;
        } else {
            ep_reward[e] = er; ep_len[e] = el;
        }
    }
    float vals[6] = {fin, rsum, lsum, win, loss, draw};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float v = vals[k];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(MNK_FULL_WARP, v, o);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&totals[k], (double)v);
    }
}

extern "C" {

int mnk_rollout_store_obs(const mnk_state_t* st, const uint8_t* agent_side, uint64_t* slot, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (slot == nullptr) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(slot) & 7u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    store_obs_kernel<<<mnk_flat_blocks(st->num_envs), kFlatThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, agent_side, reinterpret_cast<u64*>(slot));
    return mnk_launch_status();
}

int mnk_rollout_gather(int32_t m, int32_t n, int32_t k, const uint64_t* packed, int64_t num_envs, const int64_t* index,
                       int64_t count, float* obs, uint8_t* mask, void* stream) {
    if (packed == nullptr || (obs == nullptr && mask == nullptr)) return MNK_ERR_NULL;
    const int words = mnk_words_for(m, n);
    if (words < 0) return words;
    if (k < 1 || k > m || k > n) return MNK_ERR_GEOM;
    if (count < 0 || num_envs <= 0) return MNK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(packed) & 7u) || (reinterpret_cast<uintptr_t>(obs) & 7u)) return MNK_ERR_ALIGN;
    if (count == 0) return MNK_OK;
    mnk_state_t geom{m, n, k, words, num_envs, nullptr, nullptr};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(geom, [&](auto g) {
        using G = decltype(g);
        rollout_gather_kernel<<<mnk_cta_tiles(count), tile_cta_threads<G>(), 0, s>>>(g, reinterpret_cast<const u64*>(packed), num_envs, index, count, obs, mask);
        return mnk_launch_status();
    });
}

int mnk_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, int64_t steps,
            int64_t num_envs, double gamma, double gae_lambda, float* advantages, float* returns, void* stream) {
    if (!rewards || !values || !dones || !last_values || !advantages || !returns) return MNK_ERR_NULL;
    if (steps < 0 || num_envs < 0) return MNK_ERR_ARG;
    if (steps == 0 || num_envs == 0) return MNK_OK;
    gae_kernel<<<mnk_flat_blocks(num_envs), kFlatThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        rewards, values, dones, last_values, steps, num_envs, (float)gamma, (float)(gamma * gae_lambda), advantages, returns);
    return mnk_launch_status();
}

int mnk_episode_stats(const float* rewards, const uint8_t* dones, int64_t num_envs, float* ep_reward, float* ep_len,
                      double* totals, void* stream) {
    if (!rewards || !dones || !ep_reward || !ep_len || !totals) return MNK_ERR_NULL;
    if (num_envs < 0) return MNK_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(totals) & 7u) return MNK_ERR_ALIGN;
    if (num_envs == 0) return MNK_OK;
    episode_stats_kernel<<<mnk_flat_blocks(num_envs), kFlatThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        rewards, dones, num_envs, ep_reward, ep_len, totals);
    return mnk_launch_status();
}

}  // extern "C"
