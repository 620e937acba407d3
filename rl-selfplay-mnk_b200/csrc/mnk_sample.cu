// mnk_sample.cu -- masked categorical sampling / evaluation, one warp per row.
//
// Replaces, for the rollout path, the chain  where(mask, logits, -inf) -> all-masked rows to zeros
// -> Categorical(logits) -> sample() / log_prob() / entropy()  of the reference
// (src/alg/architectures/resnet.py:84-94, src/selfplay/policy.py:46-52, src/alg/ppo.py:97-100),
// which materialises several [B, A] temporaries and calls torch.multinomial.  Here each lane holds
// ceil(A/32) logits in registers; max / sum-exp / arg-max are warp shuffles; the draw is Gumbel-max
// (argmax(logit + G), G = -log(-log u)), which samples exactly softmax(logits).
#include "mnk_dispatch.cuh"

#include <math_constants.h>

template <int ITEMS>
__global__ void __launch_bounds__(128)
masked_sample_kernel(const float* __restrict__ logits, long long row_stride, const u8* __restrict__ mask, int num_actions,
                     long long rows, u64 seed, u64 counter, const u64* __restrict__ counter_base, long long row_offset,
                     int deterministic,
                     const int64_t* __restrict__ given, int64_t* __restrict__ actions, float* __restrict__ log_probs,
                     float* __restrict__ entropy) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (counter_base != nullptr) counter += *counter_base;   // device-resident base: lets a captured graph draw fresh numbers per replay
    const float* lrow = logits + (size_t)row * row_stride;
    const u8* mrow = mask ? mask + (size_t)row * num_actions : nullptr;

    float v[ITEMS];
    float vmax = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int a = lane + 32 * i;
        float x = -CUDART_INF_F;
        if (a < num_actions && (mrow == nullptr || mrow[a] != 0)) x = lrow[a];
        v[i] = x;
        vmax = fmaxf(vmax, x);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(MNK_FULL_WARP, vmax, o));
    if (vmax == -CUDART_INF_F) {   // all masked -> zeros, i.e. uniform over all actions (resnet.py:91-92)
        vmax = 0.0f;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) v[i] = (lane + 32 * i < num_actions) ? 0.0f : -CUDART_INF_F;
    }
    float sum = 0.0f, dot = 0.0f;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const float p = expf(v[i] - vmax);        // exp(-inf) = 0
        sum += p;
        dot += (p > 0.0f) ? p * v[i] : 0.0f;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        sum += __shfl_xor_sync(MNK_FULL_WARP, sum, o);
        dot += __shfl_xor_sync(MNK_FULL_WARP, dot, o);
    }
    const float lse = vmax + logf(sum);

    long long chosen;
    if (given != nullptr) {
        chosen = given[row];
    } else {
        float best = -CUDART_INF_F;
        int best_a = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int a = lane + 32 * i;
            float key = v[i];
            if (!deterministic && key != -CUDART_INF_F) {
                uint4 r = mnk_philox(seed, (u64)(row_offset + row), counter,
                                     MNK_STREAM_SAMPLE | ((u32)lane << 8) | ((u32)(i >> 2) << 16));
                const u32 bits = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
                const float u = ((float)(bits >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
                key += -logf(-logf(u));
            }
            if (key > best || (key == best && a < best_a && key != -CUDART_INF_F)) { best = key; best_a = a; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(MNK_FULL_WARP, best, o);
            const int oa = __shfl_xor_sync(MNK_FULL_WARP, best_a, o);
            if (ob > best || (ob == best && oa < best_a)) { best = ob; best_a = oa; }
        }
        chosen = best_a;
        if (lane == 0 && actions != nullptr) actions[row] = chosen;
    }
    if (log_probs != nullptr) {
        // logit of the chosen action lives in lane chosen%32, slot chosen/32
        float mine = -CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (lane + 32 * i == chosen) mine = v[i];
        const float lv = __shfl_sync(MNK_FULL_WARP, mine, (int)(chosen & 31));
        if (lane == 0) log_probs[row] = lv - lse;
    }
    if (entropy != nullptr && lane == 0) entropy[row] = lse - dot / sum;
}

extern "C" int mnk_masked_sample(const float* logits, int64_t row_stride, const uint8_t* mask, int32_t num_actions,
                                 int64_t rows, uint64_t seed, uint64_t counter, const uint64_t* counter_base,
                                 int64_t row_offset, int deterministic,
                                 const int64_t* given, int64_t* actions, float* log_probs, float* entropy,
                                 void* stream) {
    if (logits == nullptr || (given == nullptr && actions == nullptr)) return MNK_ERR_NULL;
    if (num_actions < 1 || num_actions > 512 || rows < 0 || row_stride < num_actions) return MNK_ERR_ARG;
    if (rows == 0) return MNK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((rows + 3) / 4);
    const int items = (num_actions + 31) / 32;
#define MNK_LAUNCH_SAMPLE(I)                                                                                          \
    masked_sample_kernel<I><<<blocks, 128, 0, s>>>(logits, row_stride, mask, num_actions, rows, seed, (u64)counter,  \
                                                   reinterpret_cast<const u64*>(counter_base), row_offset, deterministic, given,  \
                                                   actions, log_probs, entropy)
    if (items <= 1) MNK_LAUNCH_SAMPLE(1);
    else if (items <= 3) MNK_LAUNCH_SAMPLE(3);
    else if (items <= 6) MNK_LAUNCH_SAMPLE(6);
    else if (items <= 8) MNK_LAUNCH_SAMPLE(8);
    else if (items <= 12) MNK_LAUNCH_SAMPLE(12);
    else MNK_LAUNCH_SAMPLE(16);
#undef MNK_LAUNCH_SAMPLE
    return mnk_launch_status();
}
