// On-device sampler: repetition penalty, min_new_tokens, suppress range, temperature, top-k, top-p,
// argmax / categorical draw -- one CTA per sequence, no host round trip per token.
//
// Replaces mx.argmax / mx.random.categorical plus the host-side logits processors the reference stack
// runs once per sampled token (one device sync per token).  The arithmetic lives in sampler.cuh
// (shared with the persistent frame kernel).
#include "sampler.cuh"

namespace q3t {

constexpr int SAMPLE_THREADS = 1024;

struct SampleParams {
    const float* logits; int V; long long logits_stride;
    q3t_sampling sp;
    unsigned int* seen;
    const int* step; int step_stride;
    int rng_stream;
    const float* uniforms;
    int* out; long long out_stride; long long fo_stride, fo_step_stride;
    const int* forced; int* own;
    int* done;
};

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const SampleParams p) {
    __shared__ float sc[SAMPLE_MAXV];
    __shared__ float pe[SAMPLE_MAXV];
    __shared__ unsigned short cand[SAMPLE_MAXV];
    __shared__ unsigned int hist[256];
    __shared__ float redf[32];
    __shared__ int redi[32];
    __shared__ int sh_i[4];

    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int V = p.V;
    const int step = p.step ? p.step[(size_t)b * p.step_stride] : 0;
    const float* lg = p.logits + (size_t)b * p.logits_stride;
    const unsigned int* seen = p.seen ? p.seen + (size_t)b * ((V + 31) / 32) : nullptr;
    const q3t_sampling sp = p.sp;

    for (int i = tid; i < V; i += SAMPLE_THREADS) sc[i] = sample_score(lg[i], i, sp, seen, step);
    __syncthreads();
    SampleScratch s{sc, pe, cand, hist, redf, redi, sh_i};
    const float u = sp.do_sample ? (p.uniforms ? p.uniforms[b] : hash_uniform(sp.seed, step, p.rng_stream, b)) : 0.f;
    int choice = sample_core<SAMPLE_THREADS>(s, V, sp, u, tid, [] { __syncthreads(); });

    if (tid == 0) {
        const long long o = (long long)b * p.fo_stride + (long long)step * p.fo_step_stride;
        if (p.own) p.own[o] = choice;
        if (p.forced) choice = p.forced[o];
        p.out[(long long)b * p.out_stride] = choice;
        if (p.done && choice == sp.eos_id) p.done[b] = 1;
        if (p.seen) atomicOr(p.seen + (size_t)b * ((V + 31) / 32) + (choice >> 5), 1u << (choice & 31));
    }
}

int launch_sample(const q3t_sample_args* a, cudaStream_t stream) {
    Q3T_REQUIRE(a->V > 0 && a->V <= SAMPLE_MAXV, "sample: vocabulary larger than the shared-memory buffer (4096)");
    SampleParams p;
    p.logits = a->logits; p.V = a->V; p.logits_stride = a->logits_stride; p.sp = a->sp; p.seen = a->seen;
    p.step = a->step; p.step_stride = a->step_stride; p.rng_stream = a->rng_stream; p.uniforms = a->uniforms; p.out = a->out;
    p.out_stride = a->out_stride; p.fo_stride = a->fo_stride; p.fo_step_stride = a->fo_step_stride;
    p.forced = a->forced; p.own = a->own;
    p.done = a->done;
    launch_pdl(sample_kernel, dim3(a->B), dim3(SAMPLE_THREADS), 0, stream, p);
    Q3T_CHECK_LAUNCH("sample");
    return 0;
}

}  // namespace q3t

extern "C" int q3t_sample(const q3t_sample_args* a, void* stream) {
    return q3t::launch_sample(a, (cudaStream_t)stream);
}
