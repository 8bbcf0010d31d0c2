// Device-side driver of ONE denoising step of the guided loop: refinement loop, threshold tests,
// recursion rounds and re-noising run inside a single CUDA graph with conditional nodes, so the host enqueues one
// launch per denoising step and never reads a loss back.
//
// Restates the control flow of reference pipeline_guided_attention.py:925-1053 (one denoising step incl. its recursion
// rounds), :475-581 (_perform_iterative_refinement_step) and :1074-1088 (meets_threshold) over five device programs
// that the caller captured as CUDA graphs (PyTorch is plumbing: the UNet passes inside them are cuDNN/cuBLAS work plus
// this library's attention / tail kernels):
//
//     eval      text-conditioned UNet forward + guidance tail                       -> stats_eval
//     update    the same forward with autograd + backward + latents_out = latents - step * grad   -> stats_update
//     cfg       CFG forward (batch 2) + DDIM step                                   -> latents_out
//     advance   latents <- latents_out
//     renoise   latents <- sqrt(Bt) latents + sqrt(1 - Bt) noise[n_draws++]         (reference :1046-1050)
//
//     graph:  init -> WHILE(round) { eval; decide_round; WHILE(refine) { update; advance; decide_refine }
//                                    IF(final_update) { update; advance }  IF(final_eval) { eval }
//                                    cfg; advance; decide_next; IF(renoise) { renoise } }
//
// The decide_* nodes are single-thread kernels of this file: they group the per-token unscaled losses by sub-prompt
// (pipeline :358-387), compare with the step's thresholds in double precision like the host's Python floats, and set
// the conditional handles with cudaGraphSetConditional.  Results (latents, UNet-pass counts) are identical to the
// host-driven loop; what disappears is one D2H read + one pipeline drain per loss evaluation.
#include "ga_common.cuh"
#include <new>
#include <stddef.h>

namespace ga {
namespace step {

struct Ctl {   // device-resident control block (caller-owned, GA_STEP_CTL_BYTES)
  // inputs of the current step (written by set_params_kernel)
  double thr_call, thr_cfg, thr_last;
  int32_t has_thr_call, has_thr_cfg, has_thr_last;
  int32_t check, update_cond, recurse_ok, renoise_ok, recurse_steps, max_refine;
  // state of the current step
  int32_t round, iteration, did_update, do_update;
  // lifetime counters (never reset by the driver): UNet passes by program, refinement iterations, rounds
  int32_t n_eval, n_update, n_cfg, n_refine, n_rounds, n_renoise;
};
static_assert(sizeof(Ctl) <= GA_STEP_CTL_BYTES, "GA_STEP_CTL_BYTES too small");
static_assert(offsetof(Ctl, n_eval) == GA_STEP_COUNTER_BASE * 4, "GA_STEP_COUNTER_BASE does not match the layout");

struct Tokens {
  int32_t group[GA_MAX_TOKENS];
  int32_t kind[GA_MAX_TOKENS];
  int32_t n, n_groups, avg_within;
};

// any sub-prompt's summed (or averaged) unscaled loss > thr   (reference :358-387 + :1086)
__device__ bool exceeds(const float* stats, const float* custom, const Tokens& tk, double thr) {
  for (int g = 0; g < tk.n_groups; ++g) {
    double tot = 0.0;
    int cnt = 0;
    for (int n = 0; n < tk.n; ++n)
      if (tk.group[n] == g && tk.kind[n] != GA_TOKEN_KEYWORD) ++cnt;
    if (cnt == 0) continue;
    for (int n = 0; n < tk.n; ++n)
      if (tk.group[n] == g && tk.kind[n] != GA_TOKEN_KEYWORD) {
        const double v = (double)stats[n * GA_STATS + GA_STAT_UNSCALED];
        tot = tot + (tk.avg_within ? v / (double)cnt : v);
      }
    if (tot > thr) return true;
  }
  if (custom != nullptr && (double)custom[0] > thr) return true;   // the (None, custom_loss) entry (:446-448)
  return false;
}

__global__ void set_params_kernel(Ctl* ctl, ga_step_params_t p, long long* t_dev, float* step_dev, float* ddim_dev,
                                  float* renoise_dev) {
  if (threadIdx.x != 0) return;
  ctl->thr_call = p.thr_call; ctl->thr_cfg = p.thr_cfg; ctl->thr_last = p.thr_last;
  ctl->has_thr_call = p.has_thr_call; ctl->has_thr_cfg = p.has_thr_cfg; ctl->has_thr_last = p.has_thr_last;
  ctl->check = p.check; ctl->update_cond = p.update_cond; ctl->recurse_ok = p.recurse_ok;
  ctl->renoise_ok = p.renoise_ok; ctl->recurse_steps = p.recurse_steps; ctl->max_refine = p.max_refine;
  *t_dev = p.timestep;
  *step_dev = p.step_size;
  for (int i = 0; i < 4; ++i) ddim_dev[i] = p.ddim[i];
  renoise_dev[0] = p.renoise[0];
  renoise_dev[1] = p.renoise[1];
}

__global__ void init_kernel(Ctl* ctl) {
  if (threadIdx.x != 0) return;
  ctl->round = 0;
  ctl->iteration = 0;
  ctl->did_update = 0;
  ctl->do_update = 0;
}

// after the round's first evaluation (reference :946-1004; host mirror: pipeline `_denoise_graphed`)
__global__ void decide_round_kernel(Ctl* ctl, const float* stats, const float* custom, Tokens tk,
                                    cudaGraphConditionalHandle h_refine, cudaGraphConditionalHandle h_final_update,
                                    cudaGraphConditionalHandle h_final_eval, float* refine_first) {
  if (threadIdx.x != 0) return;
  if (refine_first != nullptr) *refine_first = 1.f;   // a refinement that starts now starts with fresh optimizer state
  ctl->n_eval += 1;
  ctl->n_rounds += 1;
  bool met = true, do_update = false;
  if (ctl->check) {
    met = !(ctl->has_thr_call && exceeds(stats, custom, tk, ctl->thr_call));
    do_update = ctl->update_cond && ctl->has_thr_last && exceeds(stats, custom, tk, ctl->thr_last);
  }
  const bool refine = !met;
  ctl->iteration = 0;
  ctl->do_update = do_update;
  ctl->did_update = (refine || do_update) ? 1 : 0;
  cudaGraphSetConditional(h_refine, refine ? 1u : 0u);
  cudaGraphSetConditional(h_final_update, do_update ? 1u : 0u);
  cudaGraphSetConditional(h_final_eval, (refine && !do_update) ? 1u : 0u);
  if (do_update) ctl->n_update += 1;            // the final evaluation carries the threshold-step update (:1003)
  else if (refine) ctl->n_eval += 1;
}

// after every refinement iteration (reference :501-557: the loop re-tests the losses it computed BEFORE the update)
__global__ void decide_refine_kernel(Ctl* ctl, const float* stats, const float* custom, Tokens tk,
                                     cudaGraphConditionalHandle h_refine, float* refine_first) {
  if (threadIdx.x != 0) return;
  if (refine_first != nullptr) *refine_first = 0.f;
  ctl->iteration += 1;
  ctl->n_update += 1;
  ctl->n_refine += 1;
  bool again = ctl->has_thr_cfg && exceeds(stats, custom, tk, ctl->thr_cfg);
  if (ctl->iteration >= ctl->max_refine) again = false;
  cudaGraphSetConditional(h_refine, again ? 1u : 0u);
}

// after the round's CFG forward + DDIM step (reference :1041-1053)
__global__ void decide_next_kernel(Ctl* ctl, cudaGraphConditionalHandle h_round, cudaGraphConditionalHandle h_renoise) {
  if (threadIdx.x != 0) return;
  ctl->n_cfg += 1;
  const bool last = ctl->round == ctl->recurse_steps - 1;
  const bool stop = !ctl->recurse_ok || !ctl->did_update;
  const bool go = !stop && !last;
  ctl->round += 1;
  cudaGraphSetConditional(h_round, go ? 1u : 0u);
  const bool renoise = go && ctl->renoise_ok;
  if (renoise) ctl->n_renoise += 1;
  cudaGraphSetConditional(h_renoise, renoise ? 1u : 0u);
}

struct Driver {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  Ctl* ctl = nullptr;
  long long* t_dev = nullptr;
  float *step_dev = nullptr, *ddim_dev = nullptr, *renoise_dev = nullptr;
};

#define GA_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(GA_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));      \
  } while (0)

// Appends `child` (cloned) to `g` after `*tail` (or as a root when *tail is null) and makes it the new tail.
static int add_child(cudaGraph_t g, cudaGraphNode_t* tail, cudaGraph_t child) {
  cudaGraphNode_t n;
  GA_CUDA(cudaGraphAddChildGraphNode(&n, g, *tail ? tail : nullptr, *tail ? 1 : 0, child));
  *tail = n;
  return GA_OK;
}

static int add_kernel(cudaGraph_t g, cudaGraphNode_t* tail, void* fn, void** args) {
  cudaKernelNodeParams kp = {};
  kp.func = fn;
  kp.gridDim = dim3(1);
  kp.blockDim = dim3(32);
  kp.sharedMemBytes = 0;
  kp.kernelParams = args;
  kp.extra = nullptr;
  cudaGraphNode_t n;
  GA_CUDA(cudaGraphAddKernelNode(&n, g, *tail ? tail : nullptr, *tail ? 1 : 0, &kp));
  *tail = n;
  return GA_OK;
}

static int add_cond(cudaGraph_t g, cudaGraphNode_t* tail, cudaGraphConditionalHandle h, cudaGraphConditionalNodeType type,
                    cudaGraph_t* body) {
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = h;
  np.conditional.type = type;
  np.conditional.size = 1;
  cudaGraphNode_t n;
  GA_CUDA(cudaGraphAddNode(&n, g, *tail ? tail : nullptr, *tail ? 1 : 0, &np));
  *body = np.conditional.phGraph_out[0];
  *tail = n;
  return GA_OK;
}

static int build(Driver* d, const ga_step_programs_t* pr, const float* stats_eval, const float* stats_update,
                 const float* custom_eval, const float* custom_update, const float* stats_refine,
                 const float* custom_refine, float* refine_first, const Tokens& tk) {
  cudaGraph_t ev = (cudaGraph_t)pr->eval, up = (cudaGraph_t)pr->update, cf = (cudaGraph_t)pr->cfg,
              adv = (cudaGraph_t)pr->advance, rn = (cudaGraph_t)pr->renoise;
  cudaGraph_t rup = pr->refine_update != nullptr ? (cudaGraph_t)pr->refine_update : up;
  GA_CUDA(cudaGraphCreate(&d->graph, 0));
  cudaGraph_t G = d->graph;
  int rc;
  cudaGraphNode_t tail = nullptr;
  Ctl* ctl = d->ctl;
  {
    void* a[] = {&ctl};
    if ((rc = add_kernel(G, &tail, (void*)init_kernel, a)) != GA_OK) return rc;
  }
  // --- WHILE (round): default 1 at every launch
  cudaGraphConditionalHandle h_round;
  GA_CUDA(cudaGraphConditionalHandleCreate(&h_round, G, 1, cudaGraphCondAssignDefault));
  cudaGraph_t R;
  if ((rc = add_cond(G, &tail, h_round, cudaGraphCondTypeWhile, &R)) != GA_OK) return rc;

  cudaGraphConditionalHandle h_refine, h_fu, h_fe, h_rn;
  GA_CUDA(cudaGraphConditionalHandleCreate(&h_refine, R, 0, 0));
  GA_CUDA(cudaGraphConditionalHandleCreate(&h_fu, R, 0, 0));
  GA_CUDA(cudaGraphConditionalHandleCreate(&h_fe, R, 0, 0));
  GA_CUDA(cudaGraphConditionalHandleCreate(&h_rn, R, 0, 0));
  cudaGraphNode_t rt = nullptr;
  if ((rc = add_child(R, &rt, ev)) != GA_OK) return rc;
  {
    Tokens t = tk;
    void* a[] = {&ctl, (void*)&stats_eval, (void*)&custom_eval, &t, &h_refine, &h_fu, &h_fe, &refine_first};
    if ((rc = add_kernel(R, &rt, (void*)decide_round_kernel, a)) != GA_OK) return rc;
  }
  {   // WHILE (refine) { update; advance; decide_refine }
    cudaGraph_t F;
    if ((rc = add_cond(R, &rt, h_refine, cudaGraphCondTypeWhile, &F)) != GA_OK) return rc;
    cudaGraphNode_t ft = nullptr;
    if ((rc = add_child(F, &ft, rup)) != GA_OK) return rc;
    if ((rc = add_child(F, &ft, adv)) != GA_OK) return rc;
    Tokens t = tk;
    void* a[] = {&ctl, (void*)&stats_refine, (void*)&custom_refine, &t, &h_refine, &refine_first};
    if ((rc = add_kernel(F, &ft, (void*)decide_refine_kernel, a)) != GA_OK) return rc;
  }
  {   // IF (final_update) { update; advance }
    cudaGraph_t U;
    if ((rc = add_cond(R, &rt, h_fu, cudaGraphCondTypeIf, &U)) != GA_OK) return rc;
    cudaGraphNode_t ut = nullptr;
    if ((rc = add_child(U, &ut, up)) != GA_OK) return rc;
    if ((rc = add_child(U, &ut, adv)) != GA_OK) return rc;
  }
  {   // IF (final_eval) { eval }
    cudaGraph_t E;
    if ((rc = add_cond(R, &rt, h_fe, cudaGraphCondTypeIf, &E)) != GA_OK) return rc;
    cudaGraphNode_t et = nullptr;
    if ((rc = add_child(E, &et, ev)) != GA_OK) return rc;
  }
  if ((rc = add_child(R, &rt, cf)) != GA_OK) return rc;
  if ((rc = add_child(R, &rt, adv)) != GA_OK) return rc;
  {
    void* a[] = {&ctl, &h_round, &h_rn};
    if ((rc = add_kernel(R, &rt, (void*)decide_next_kernel, a)) != GA_OK) return rc;
  }
  {   // IF (renoise) { renoise }
    cudaGraph_t N;
    if ((rc = add_cond(R, &rt, h_rn, cudaGraphCondTypeIf, &N)) != GA_OK) return rc;
    cudaGraphNode_t nt = nullptr;
    if ((rc = add_child(N, &nt, rn)) != GA_OK) return rc;
  }
  GA_CUDA(cudaGraphInstantiate(&d->exec, G, 0));
  return GA_OK;
}

}  // namespace step
}  // namespace ga

using namespace ga;

extern "C" int ga_step_driver_create(void** driver_out, const ga_step_programs_t* programs, void* ctl_dev,
                                     const float* stats_eval, const float* stats_update, const float* custom_eval,
                                     const float* custom_update, const float* stats_refine,
                                     const float* custom_refine, float* refine_first_dev,
                                     const ga_token_t* tokens_host, int n_tokens,
                                     int n_groups, int avg_within, int64_t* t_dev, float* step_dev, float* ddim_dev,
                                     float* renoise_dev) {
  GA_CHECK_ARG(driver_out != nullptr && programs != nullptr && ctl_dev != nullptr, "NULL argument");
  GA_CHECK_ARG(programs->eval && programs->update && programs->cfg && programs->advance && programs->renoise,
               "all five device programs are required");
  GA_CHECK_ARG(stats_eval != nullptr && stats_update != nullptr, "stats pointers are required");
  GA_CHECK_ARG(programs->refine_update == nullptr || stats_refine != nullptr,
               "a separate refinement program needs its own stats buffer");
  if (programs->refine_update == nullptr) { stats_refine = stats_update; custom_refine = custom_update; }
  GA_CHECK_ARG(n_tokens >= 0 && n_tokens <= GA_MAX_TOKENS && (n_tokens == 0 || tokens_host != nullptr),
               "n_tokens %d out of range", n_tokens);
  GA_CHECK_ARG(t_dev && step_dev && ddim_dev && renoise_dev, "per-step scalar buffers are required");
  step::Tokens tk = {};
  tk.n = n_tokens;
  tk.n_groups = n_groups;
  tk.avg_within = avg_within;
  for (int i = 0; i < n_tokens; ++i) {
    tk.group[i] = tokens_host[i].group;
    tk.kind[i] = tokens_host[i].kind;
  }
  step::Driver* d = new (std::nothrow) step::Driver();
  if (d == nullptr) return fail(GA_ERR_CUDA, "out of host memory");
  d->ctl = static_cast<step::Ctl*>(ctl_dev);
  d->t_dev = reinterpret_cast<long long*>(t_dev);
  d->step_dev = step_dev;
  d->ddim_dev = ddim_dev;
  d->renoise_dev = renoise_dev;
  int rc = step::build(d, programs, stats_eval, stats_update, custom_eval, custom_update, stats_refine, custom_refine,
                       refine_first_dev, tk);
  if (rc != GA_OK) {
    if (d->exec) cudaGraphExecDestroy(d->exec);
    if (d->graph) cudaGraphDestroy(d->graph);
    delete d;
    cudaGetLastError();
    return rc;
  }
  *driver_out = d;
  return GA_OK;
}

extern "C" int ga_step_driver_run(void* driver, const ga_step_params_t* params_host, ga_stream_t stream) {
  GA_CHECK_ARG(driver != nullptr && params_host != nullptr, "NULL argument");
  step::Driver* d = static_cast<step::Driver*>(driver);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  step::set_params_kernel<<<1, 32, 0, st>>>(d->ctl, *params_host, d->t_dev, d->step_dev, d->ddim_dev, d->renoise_dev);
  int rc = check_launch("step set_params");
  if (rc != GA_OK) return rc;
  cudaError_t e = cudaGraphLaunch(d->exec, st);
  if (e != cudaSuccess) return fail(GA_ERR_CUDA, "cudaGraphLaunch(step driver): %s", cudaGetErrorString(e));
  return GA_OK;
}

extern "C" int ga_step_driver_set_params(void* driver, const ga_step_params_t* params_host, ga_stream_t stream) {
  GA_CHECK_ARG(driver != nullptr && params_host != nullptr, "NULL argument");
  step::Driver* d = static_cast<step::Driver*>(driver);
  step::set_params_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(d->ctl, *params_host, d->t_dev, d->step_dev,
                                                                             d->ddim_dev, d->renoise_dev);
  return check_launch("step set_params");
}

extern "C" int ga_step_driver_destroy(void* driver) {
  if (driver == nullptr) return GA_OK;
  step::Driver* d = static_cast<step::Driver*>(driver);
  if (d->exec) cudaGraphExecDestroy(d->exec);
  if (d->graph) cudaGraphDestroy(d->graph);
  delete d;
  return GA_OK;
}
