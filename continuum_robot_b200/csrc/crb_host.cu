// crb_host.cu -- host-resident ensembles: chunked H2D / fused RK4 / D2H pipeline (host code only).
//
// The reference keeps every beam's state in host memory and fans whole simulations out over
// processes (examples/beam_comparison_gravity.py:72-73, examples/example_utilities.py:116-170).  A
// caller that keeps the ensemble state in (pinned) host memory gets the same contract here:
// crb_rk4_host advances X_host[B,2n] in place.  The ensemble is cut into member chunks of whole
// kernel waves (default two); chunk c+1 is copied in and chunk c-1 copied out while chunk c integrates
// (three streams, PCIe is full duplex), and consecutive calls on the same buffers overlap chunk
// by chunk (the only cross-call ordering is "chunk c is not uploaded before its previous result
// has been downloaded"), so a sequence of calls runs at max(copy-in, compute, copy-out).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "crb_internal.h"

struct crb_pipeline {
  int device = -1;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t ev_entry = nullptr, ev_all_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_cmp, ev_out;
  // layout of the previous call (per-chunk cross-call dependencies hold only if it repeats)
  const double* last_host = nullptr;
  const double* last_dev = nullptr;
  long long last_members = 0;
  int last_chunk = 0, last_row = 0;
  int last_trace_chunks = 0;
  bool pending = false;
};

#define CRB_CUDA(call, who)                                                                        \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return crb_fail(CRB_E_CUDA, "%s: %s: %s", who, #call, cudaGetErrorString(e_)); \
  } while (0)

static bool pipeline_trace() {  // CRB_PIPELINE_TRACE=1: events carry time stamps, crb_pipeline_synchronize prints the chunk timeline of the last call
  static const bool on = getenv("CRB_PIPELINE_TRACE") != nullptr;
  return on;
}

static int ensure_events(crb_pipeline* p, size_t n) {
  const unsigned flags = pipeline_trace() ? cudaEventDefault : cudaEventDisableTiming;
  while (p->ev_in.size() < n) {
    cudaEvent_t a, b, c;
    CRB_CUDA(cudaEventCreateWithFlags(&a, flags), "crb_pipeline");
    CRB_CUDA(cudaEventCreateWithFlags(&b, flags), "crb_pipeline");
    CRB_CUDA(cudaEventCreateWithFlags(&c, flags), "crb_pipeline");
    p->ev_in.push_back(a);
    p->ev_cmp.push_back(b);
    p->ev_out.push_back(c);
  }
  return 0;
}

extern "C" int crb_pipeline_destroy(crb_pipeline_t* p);

static int pipeline_build(crb_pipeline* p) {
  CRB_CUDA(cudaGetDevice(&p->device), "crb_pipeline_create");
  CRB_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking), "crb_pipeline_create");
  CRB_CUDA(cudaStreamCreateWithFlags(&p->s_cmp, cudaStreamNonBlocking), "crb_pipeline_create");
  CRB_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking), "crb_pipeline_create");
  CRB_CUDA(cudaEventCreateWithFlags(&p->ev_entry, pipeline_trace() ? cudaEventDefault : cudaEventDisableTiming), "crb_pipeline_create");
  CRB_CUDA(cudaEventCreateWithFlags(&p->ev_all_out, cudaEventDisableTiming), "crb_pipeline_create");
  return 0;
}

extern "C" int crb_pipeline_create(crb_pipeline_t** out) {
  if (!out) return crb_fail(CRB_E_ARG, "crb_pipeline_create: null output");
  crb_pipeline* p = new crb_pipeline();
  if (int rc = pipeline_build(p)) {  // release whatever was created before the failure (the error text is kept)
    crb_pipeline_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}

extern "C" int crb_pipeline_destroy(crb_pipeline_t* p) {
  if (!p) return 0;
  for (cudaStream_t st : {p->s_in, p->s_cmp, p->s_out})
    if (st) cudaStreamSynchronize(st);
  for (auto v : {&p->ev_in, &p->ev_cmp, &p->ev_out})
    for (cudaEvent_t e : *v) cudaEventDestroy(e);
  if (p->ev_entry) cudaEventDestroy(p->ev_entry);
  if (p->ev_all_out) cudaEventDestroy(p->ev_all_out);
  for (cudaStream_t st : {p->s_in, p->s_cmp, p->s_out})
    if (st) cudaStreamDestroy(st);
  delete p;
  return 0;
}

extern "C" int crb_pipeline_wait(crb_pipeline_t* p, void* stream) {
  if (!p) return crb_fail(CRB_E_ARG, "crb_pipeline_wait: null pipeline");
  if (!p->pending) return 0;
  CRB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, p->ev_all_out, 0), "crb_pipeline_wait");
  return 0;
}

extern "C" int crb_pipeline_synchronize(crb_pipeline_t* p) {
  if (!p) return crb_fail(CRB_E_ARG, "crb_pipeline_synchronize: null pipeline");
  CRB_CUDA(cudaStreamSynchronize(p->s_in), "crb_pipeline_synchronize");
  CRB_CUDA(cudaStreamSynchronize(p->s_cmp), "crb_pipeline_synchronize");
  CRB_CUDA(cudaStreamSynchronize(p->s_out), "crb_pipeline_synchronize");
  if (pipeline_trace() && p->pending && p->last_trace_chunks > 0) {
    fprintf(stderr, "[crb pipeline] chunk: copy-in done / kernel done / copy-out done (ms after the call's entry)\n");
    for (int c = 0; c < p->last_trace_chunks; ++c) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, p->ev_entry, p->ev_in[c]);
      cudaEventElapsedTime(&b, p->ev_entry, p->ev_cmp[c]);
      cudaEventElapsedTime(&d, p->ev_entry, p->ev_out[c]);
      fprintf(stderr, "[crb pipeline] %2d: %.3f %.3f %.3f\n", c, a, b, d);
    }
  }
  p->pending = false;
  return 0;
}

extern "C" int crb_system_slice(const crb_plan_t* plan, const crb_system_t* sys, int32_t lo, int32_t count,
                                crb_system_t* out) {
  if (!plan || !sys || !out) return crb_fail(CRB_E_ARG, "crb_system_slice: null argument");
  if (lo < 0 || count < 1 || (long long)lo + count > sys->n_members)
    return crb_fail(CRB_E_ARG, "crb_system_slice: members [%d, %d) outside the %d-member system", lo, lo + count,
                    sys->n_members);
  *out = *sys;
  out->n_members = count;
  const long long n = plan->n_free, P = plan->p, N = plan->n_elements;
  if (!sys->mass_shared && sys->mfac) out->mfac = sys->mfac + lo * plan->mfac_doubles;
  if (!sys->stiff_shared && sys->kcoef) out->kcoef = sys->kcoef + lo * P * 4;
  if (!sys->force_shared) {
    if (sys->drag) out->drag = sys->drag + lo * P;
    if (sys->grav) out->grav = sys->grav + lo * P * 2;
    if (sys->seg_half_mass) out->seg_half_mass = sys->seg_half_mass + lo * N;
  }
  if (sys->u_const) out->u_const = sys->u_const + lo * n;
  if (sys->f_ext) out->f_ext = sys->f_ext + lo * n;
  if (sys->imp_amp) out->imp_amp = sys->imp_amp + lo;
  if (!sys->u_time_shared) {
    if (sys->u_sin_amp) out->u_sin_amp = sys->u_sin_amp + lo * n;
    if (sys->u_tab_v) return crb_fail(CRB_E_ARG, "crb_system_slice: a per-member input table [K,B,n] cannot be sliced by member (its knot stride is the full ensemble)");
  }
  if (sys->gain && sys->gain_stride) out->gain = sys->gain + lo * sys->gain_stride;
  if (sys->member_op) out->member_op = sys->member_op + lo * n * (3 * n + 1);
  out->member_order = nullptr;  // a launch order of the whole ensemble does not carry over to a sub-range
  return 0;
}

extern "C" int crb_rk4_host(crb_pipeline_t* p, const crb_plan_t* plan, const crb_system_t* sys, double* X_host,
                            double* X_dev, int32_t chunk_members, double t0, double h, int32_t nsteps,
                            void* stream) {
  if (!p || !plan || !sys) return crb_fail(CRB_E_ARG, "crb_rk4_host: null pipeline/plan/system");
  if (!X_host || !X_dev) return crb_fail(CRB_E_ARG, "crb_rk4_host: null state pointer");
  if (nsteps < 0) return crb_fail(CRB_E_ARG, "crb_rk4_host: nsteps must be >= 0");
  if (nsteps == 0) return 0;
  int dev = -1;
  CRB_CUDA(cudaGetDevice(&dev), "crb_rk4_host");
  if (dev != p->device) return crb_fail(CRB_E_ARG, "crb_rk4_host: pipeline was created on device %d, current is %d", p->device, dev);
  const long long B = sys->n_members;
  int chunk = chunk_members;
  if (chunk <= 0) {
    if (int rc = crb_rk4_wave_members(plan, sys, &chunk)) return rc;
    chunk *= 2;  // two waves per chunk: measured best on B200 (fewer, larger copies; 2.11 vs 2.21 ms per 100 MB call)
  }
  if (chunk > B) chunk = (int)B;
  const size_t nchunks = (size_t)((B + chunk - 1) / chunk);
  if (int rc = ensure_events(p, nchunks)) return rc;
  const int row = 2 * plan->n_free;
  // chunk-wise ordering against the previous call is valid only if that call used the same layout
  const bool same = p->pending && p->last_host == X_host && p->last_dev == X_dev && p->last_members == B &&
                    p->last_chunk == chunk && p->last_row == row;
  // everything the caller enqueued on its stream so far precedes this call's first copy
  CRB_CUDA(cudaEventRecord(p->ev_entry, (cudaStream_t)stream), "crb_rk4_host");
  CRB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_entry, 0), "crb_rk4_host");
  if (p->pending && !same) CRB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_all_out, 0), "crb_rk4_host");
  auto enqueue_chunk = [&](size_t c) -> int {
    const long long lo = (long long)c * chunk;
    const int cnt = (int)((lo + chunk <= B) ? chunk : B - lo);
    const size_t bytes = sizeof(double) * (size_t)cnt * row;
    double* xd = X_dev + lo * row;
    double* xh = X_host + lo * row;
    if (same) CRB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_out[c], 0), "crb_rk4_host");  // previous result of this chunk is on the host
    CRB_CUDA(cudaMemcpyAsync(xd, xh, bytes, cudaMemcpyHostToDevice, p->s_in), "crb_rk4_host");
    CRB_CUDA(cudaEventRecord(p->ev_in[c], p->s_in), "crb_rk4_host");
    CRB_CUDA(cudaStreamWaitEvent(p->s_cmp, p->ev_in[c], 0), "crb_rk4_host");
    crb_system_t part;
    if (int rc = crb_system_slice(plan, sys, (int32_t)lo, cnt, &part)) return rc;
    if (int rc = crb_rk4(plan, &part, xd, t0, h, nsteps, nullptr, 0, p->s_cmp)) return rc;
    CRB_CUDA(cudaEventRecord(p->ev_cmp[c], p->s_cmp), "crb_rk4_host");
    CRB_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[c], 0), "crb_rk4_host");
    CRB_CUDA(cudaMemcpyAsync(xh, xd, bytes, cudaMemcpyDeviceToHost, p->s_out), "crb_rk4_host");
    CRB_CUDA(cudaEventRecord(p->ev_out[c], p->s_out), "crb_rk4_host");
    return 0;
  };
  for (size_t c = 0; c < nchunks; ++c) {
    if (int rc = enqueue_chunk(c)) {
      // chunks < c are already in flight: close the call so that crb_pipeline_wait covers them, and forget the layout
      // so that the next call takes the full-wait path instead of the chunk-wise cross-call ordering
      cudaStreamWaitEvent(p->s_out, p->ev_entry, 0);
      cudaStreamWaitEvent(p->s_out, p->ev_in[c], 0);  // may be stale: harmless, it only adds ordering
      for (cudaStream_t st : {p->s_in, p->s_cmp}) {
        cudaEvent_t tail = p->ev_cmp[c];
        if (cudaEventRecord(tail, st) == cudaSuccess) cudaStreamWaitEvent(p->s_out, tail, 0);
      }
      cudaEventRecord(p->ev_all_out, p->s_out);
      p->last_host = p->last_dev = nullptr;
      p->last_members = 0;
      p->pending = true;
      return rc;
    }
  }
  CRB_CUDA(cudaEventRecord(p->ev_all_out, p->s_out), "crb_rk4_host");
  p->last_host = X_host;
  p->last_dev = X_dev;
  p->last_members = B;
  p->last_chunk = chunk;
  p->last_row = row;
  p->last_trace_chunks = (int)nchunks;
  p->pending = true;
  return 0;
}
