// mnk_selfplay.cu -- the self-play wrapper's step as fused kernels.
//
// Replaces TorchSelfPlayWrapper.reset/step/_opponent_move_if_needed/_get_canonical_obs of the
// reference (src/selfplay/torch_self_play_wrapper.py:19-112).  The reference walks index lists
// (nonzero / boolean-mask indexing, ~12-20 host syncs and 5-7 full board clones per step); here one
// wrapper step is dense and predicated over all envs with no host synchronisation:
//
//   agent phase    (:33-56)  envs flagged pending are reset and get a new side, every other env plays
//                            the agent's action; emits which envs the opponent must answer and,
//                            for a Python opponent, the opponent's canonical view (:83-94)
//   opponent phase (:58-67)  applies the opponent's actions, folds its result into reward /
//                            terminated (reward -= r_opp, :62), sets pending = terminated (:65) and
//                            materialises the agent's canonical view with the all-masked fix (:99-112)
//   fused random   both phases in one launch when the opponent is RandomPolicy (policy.py:13-29)
//
// Same CTA-per-tile shape as the env kernels (mnk_env.cu).
#include "mnk_dispatch.cuh"

enum : u32 { OPP_IDLE = 0, OPP_PLAYS = 1, OPP_OPENS = 2 };   // OPENS: freshly reset env, result discarded (:46)

struct AgentOut {
    float reward;
    bool terminated;
    u32 opp;   // OPP_*
};

// wrapper.step for one env up to (not including) the opponent's answer
template <class G>
MNK_DEV AgentOut agent_phase(const G& g, EnvRegs<G>& s, const mnk_selfplay_t& sp, long long e, bool pending,
                             long long action, const int64_t* __restrict__ forced_sides, u32& side) {
    AgentOut out;
    if (pending) {                                         // :39-46
        env_zero(s);
        const u32 episode = sp.episodes[e] + 1u;
        sp.episodes[e] = episode;
        side = forced_sides ? (u32)(forced_sides[e] & 1)
                            : (mnk_philox(sp.seed, (u64)(sp.env_offset + e), episode, MNK_STREAM_SIDE).x & 1u);
        sp.agent_side[e] = (u8)side;
        out.reward = 0.0f;
        out.terminated = false;
        out.opp = ((s.meta & 1u) != side) ? OPP_OPENS : OPP_IDLE;   // black to move: opponent opens iff agent is white
    } else {                                               // :48-56
        const MoveResult r = apply_move(g, s, action);
        out.reward = r.reward;
        out.terminated = r.done;
        out.opp = (!r.done && (s.meta & 1u) != side) ? OPP_PLAYS : OPP_IDLE;   // :74-81
    }
    return out;
}

// the opponent's ply and its effect on the agent's transition (:58-65)
template <class G>
MNK_DEV void opponent_phase(const G& g, EnvRegs<G>& s, u32 opp, long long opp_action, float& reward, bool& terminated) {
    if (opp == OPP_IDLE) return;
    const MoveResult r = apply_move(g, s, opp_action);
    if (opp == OPP_PLAYS) {
        reward -= r.reward;        // :62
        terminated = r.done;       // :63
    }
}

template <bool ACT32>
MNK_DEV long long read_action(const void* actions, long long e) {
    return ACT32 ? (long long)static_cast<const int32_t*>(actions)[e] : (long long)static_cast<const int64_t*>(actions)[e];
}

// ------------------------------------------------------------------------------------------------
// agent phase kernel
// ------------------------------------------------------------------------------------------------
template <class G, bool ACT32>
__global__ void __launch_bounds__(tile_cta_threads<G>())
selfplay_agent_kernel(G g, mnk_state_t st, mnk_selfplay_t sp, const void* __restrict__ actions,
                      const int64_t* __restrict__ forced_sides, float* __restrict__ rewards,
                      u8* __restrict__ terminated, u8* __restrict__ opp_active, float* __restrict__ opp_obs,
                      u8* __restrict__ opp_mask, u32 flags) {
    __shared__ u32 tile_smem[TileStream<G>::kWords];
    const int lane = threadIdx.x & 31;
    const long long e0 = (long long)blockIdx.x * kTileEnvs;
    const int tile_envs = (int)min((long long)kTileEnvs, st.num_envs - e0);
    const bool emit = opp_obs != nullptr || opp_mask != nullptr;
    const bool stream = emit && tile_streams<G>(tile_envs, opp_obs, opp_mask);
    u64 obsd[G::NWD];
    u64 legd[G::NWL];
    if (threadIdx.x < 32) {
        const long long e = e0 + lane;
        EnvRegs<G> s;
        env_zero(s);
        if (e < st.num_envs) {
            env_load(st, e, s);
            u32 side = sp.agent_side[e];
            const bool pending = (flags & MNK_SP_RESET_ALL) || sp.pending[e] != 0;
            const long long a = (pending || actions == nullptr) ? 0 : read_action<ACT32>(actions, e);
            const AgentOut out = agent_phase(g, s, sp, e, pending, a, forced_sides, side);
            rewards[e] = out.reward;
            terminated[e] = out.terminated ? 1 : 0;
            opp_active[e] = (u8)out.opp;
            env_store(st, e, s);
        }
        // the side to move sees its own stones on channel 0 (:87-89); raw legal mask
        if (emit) build_views(g, s, (s.meta & 1u) != 0u, false, obsd, legd);
        if (emit && !stream) emit_tile(g, e0, tile_envs, lane, obsd, legd, opp_obs, opp_mask);
    }
    if (stream) emit_block_stream_any(g, tile_smem, e0, obsd, legd, opp_obs, opp_mask);
}

// ------------------------------------------------------------------------------------------------
// opponent phase kernel (actions supplied) and fused agent + random-opponent kernel
// ------------------------------------------------------------------------------------------------
template <class G, bool ACT32, bool FUSED_RANDOM>
__global__ void __launch_bounds__(tile_cta_threads<G>())
selfplay_finish_kernel(G g, mnk_state_t st, mnk_selfplay_t sp, const void* __restrict__ agent_actions,
                       const int64_t* __restrict__ forced_sides, const void* __restrict__ opp_actions,
                       const u8* __restrict__ opp_active, u64 step_counter, float* __restrict__ rewards,
                       u8* __restrict__ terminated, float* __restrict__ obs, u8* __restrict__ mask, u32 flags) {
    __shared__ u32 tile_smem[TileStream<G>::kWords];
    const int lane = threadIdx.x & 31;
    const long long e0 = (long long)blockIdx.x * kTileEnvs;
    const int tile_envs = (int)min((long long)kTileEnvs, st.num_envs - e0);
    const bool emit = obs != nullptr || mask != nullptr;
    const bool stream = emit && tile_streams<G>(tile_envs, obs, mask);
    u64 obsd[G::NWD];
    u64 legd[G::NWL];
    if (threadIdx.x < 32) {
        const long long e = e0 + lane;
        EnvRegs<G> s;
        env_zero(s);
        u32 side = 0;
        if (e < st.num_envs) {
            env_load(st, e, s);
            side = sp.agent_side[e];
            float reward;
            bool term;
            u32 opp;
            long long oa = 0;
            if constexpr (FUSED_RANDOM) {
                const bool pending = (flags & MNK_SP_RESET_ALL) || sp.pending[e] != 0;
                const long long a = (pending || agent_actions == nullptr) ? 0 : read_action<ACT32>(agent_actions, e);
                const AgentOut out = agent_phase(g, s, sp, e, pending, a, forced_sides, side);
                reward = out.reward;
                term = out.terminated;
                opp = out.opp;
                if (opp != OPP_IDLE) {
                    const bool det = (flags & MNK_SP_DETERMINISTIC_OPP) != 0;
                    const u64 ctr = step_counter + (sp.counter_base ? *sp.counter_base : 0ull);
                    const u32 rnd = det ? 0u : mnk_philox(sp.seed, (u64)(sp.env_offset + e), ctr, MNK_STREAM_OPPONENT).x;
                    oa = pick_legal(g, s, rnd, det);
                }
            } else {
                reward = rewards[e];
                term = terminated[e] != 0;
                opp = opp_active[e];
                if (opp != OPP_IDLE) oa = read_action<ACT32>(opp_actions, e);
            }
            opponent_phase(g, s, opp, oa, reward, term);
            rewards[e] = reward;
            terminated[e] = term ? 1 : 0;
            sp.pending[e] = term ? 1 : 0;            // :65
            env_store(st, e, s);
        }
        // agent's canonical view: own stones first, all-masked rows get cell 0 (:99-112)
        if (emit) build_views(g, s, side != 0u, true, obsd, legd);
        if (emit && !stream) emit_tile(g, e0, tile_envs, lane, obsd, legd, obs, mask);
    }
    if (stream) emit_block_stream_any(g, tile_smem, e0, obsd, legd, obs, mask);
}

// ================================================================================================
// C ABI
// ================================================================================================
static int check_selfplay(const mnk_state_t* st, const mnk_selfplay_t* sp) {
    if (int rc = mnk_check_state(st)) return rc;
    if (sp == nullptr || sp->agent_side == nullptr || sp->pending == nullptr || sp->episodes == nullptr) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(sp->episodes) & 3u) return MNK_ERR_ALIGN;
    return MNK_OK;
}

extern "C" {

int mnk_selfplay_agent(const mnk_state_t* st, const mnk_selfplay_t* sp, const void* actions,
                       const int64_t* forced_sides, float* rewards, uint8_t* terminated, uint8_t* opp_active,
                       float* opp_obs, uint8_t* opp_mask, uint32_t flags, void* stream) {
    if (int rc = check_selfplay(st, sp)) return rc;
    if (rewards == nullptr || terminated == nullptr || opp_active == nullptr) return MNK_ERR_NULL;
    if (actions == nullptr && !(flags & MNK_SP_RESET_ALL)) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(opp_obs) & 7u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(*st, [&](auto g) {
        using G = decltype(g);
        const unsigned blocks = mnk_cta_tiles(st->num_envs);
        const int threads = (opp_obs != nullptr || opp_mask != nullptr) ? tile_cta_threads<G>() : 32;
        if (flags & MNK_SP_ACTIONS_I32)
            selfplay_agent_kernel<G, true><<<blocks, threads, 0, s>>>(g, *st, *sp, actions, forced_sides, rewards, terminated,
                                                                      opp_active, opp_obs, opp_mask, flags);
        else
            selfplay_agent_kernel<G, false><<<blocks, threads, 0, s>>>(g, *st, *sp, actions, forced_sides, rewards, terminated,
                                                                       opp_active, opp_obs, opp_mask, flags);
        return mnk_launch_status();
    });
}

int mnk_selfplay_opponent(const mnk_state_t* st, const mnk_selfplay_t* sp, const void* opp_actions,
                          const uint8_t* opp_active, float* rewards, uint8_t* terminated, float* obs, uint8_t* mask,
                          uint32_t flags, void* stream) {
    if (int rc = check_selfplay(st, sp)) return rc;
    if (opp_actions == nullptr || opp_active == nullptr || rewards == nullptr || terminated == nullptr) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(obs) & 7u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(*st, [&](auto g) {
        using G = decltype(g);
        const unsigned blocks = mnk_cta_tiles(st->num_envs);
        const int threads = (obs != nullptr || mask != nullptr) ? tile_cta_threads<G>() : 32;
        if (flags & MNK_SP_ACTIONS_I32)
            selfplay_finish_kernel<G, true, false><<<blocks, threads, 0, s>>>(g, *st, *sp, nullptr, nullptr, opp_actions, opp_active,
                                                                              0ull, rewards, terminated, obs, mask, flags);
        else
            selfplay_finish_kernel<G, false, false><<<blocks, threads, 0, s>>>(g, *st, *sp, nullptr, nullptr, opp_actions, opp_active,
                                                                               0ull, rewards, terminated, obs, mask, flags);
        return mnk_launch_status();
    });
}

int mnk_selfplay_step_random(const mnk_state_t* st, const mnk_selfplay_t* sp, const void* actions,
                             const int64_t* forced_sides, uint64_t step_counter, float* rewards, uint8_t* terminated,
                             float* obs, uint8_t* mask, uint32_t flags, void* stream) {
    if (int rc = check_selfplay(st, sp)) return rc;
    if (rewards == nullptr || terminated == nullptr) return MNK_ERR_NULL;
    if (actions == nullptr && !(flags & MNK_SP_RESET_ALL)) return MNK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(obs) & 7u) return MNK_ERR_ALIGN;
    if (st->num_envs == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return mnk_dispatch_geom(*st, [&](auto g) {
        using G = decltype(g);
        const unsigned blocks = mnk_cta_tiles(st->num_envs);
        const int threads = (obs != nullptr || mask != nullptr) ? tile_cta_threads<G>() : 32;
        if (flags & MNK_SP_ACTIONS_I32)
            selfplay_finish_kernel<G, true, true><<<blocks, threads, 0, s>>>(g, *st, *sp, actions, forced_sides, nullptr, nullptr,
                                                                             (u64)step_counter, rewards, terminated, obs, mask, flags);
        else
            selfplay_finish_kernel<G, false, true><<<blocks, threads, 0, s>>>(g, *st, *sp, actions, forced_sides, nullptr, nullptr,
                                                                              (u64)step_counter, rewards, terminated, obs, mask, flags);
        return mnk_launch_status();
    });
}

}  // extern "C"
