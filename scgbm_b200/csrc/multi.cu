// multi.cu -- in-process multi-GPU driver: one gbm.sc() call from ONE caller (the reference's only calling
// convention, R/fit.R:52) uses `ngpu` GPUs of the box.  The cells (columns of Y) are split into contiguous
// shards, one host thread and one NCCL rank per GPU (ncclCommInitRank on a shared unique id from ngpu threads
// of this process == ncclCommInitAll), and every rank runs the very same FitRun the one-process-per-GPU path
// runs (fit.cu) -- the collectives are the ones listed in SURVEY 8e.  Rank 0 runs on the CALLING thread, so the
// progress callback (Rprintf / R_CheckUserInterrupt in the R shim) fires on the caller's thread as scgbm.h
// promises; the worker threads never call back.
#include <thread>
#include "common.cuh"

namespace scgbm {

Comm* comm_create(const uint8_t* id, int nranks, int rank, int device);
void comm_unique_id(uint8_t* id);
void comm_destroy(Comm* c);
struct InprocGroup;
InprocGroup* inproc_group_create(int n);
void inproc_group_destroy(InprocGroup* g);
void inproc_group_abort(InprocGroup* g);
void inproc_group_finish(InprocGroup* g);
void comm_set_in_process(Comm* c, InprocGroup* g);
void fit_impl(const scgbm_fit_args* a, scgbm_fit_out* out);
void get_se_impl(const scgbm_se_args* a, scgbm_se_out* out);

static int dtype_size(int dt) {
  switch (dt) {
    case SCGBM_I32: case SCGBM_F32: return 4;
    case SCGBM_F64: return 8;
    case SCGBM_U16: return 2;
    case SCGBM_U8: return 1;
    default: throw Error(SCGBM_ERR_INVALID, "unknown dtype");
  }
}

static void shard(int64_t J, int r, int n, int64_t& j0, int64_t& j1) {   // same split as scgbm_b200/dist.py shard_range
  const int64_t base = J / n, rem = J % n;
  j0 = r * base + std::min<int64_t>(r, rem);
  j1 = j0 + base + (r < rem ? 1 : 0);
}

static std::vector<int> pick_devices(int ngpu, const int32_t* devices, int first) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(SCGBM_ERR_CUDA, "no CUDA device available: libscgbm has no CPU fallback");
  }
  if (ngpu > ndev) {
    char b[128];
    snprintf(b, sizeof(b), "ngpu = %d but only %d CUDA device(s) are visible", ngpu, ndev);
    throw Error(SCGBM_ERR_INVALID, b);
  }
  std::vector<int> dv(ngpu);
  for (int r = 0; r < ngpu; ++r) {
    dv[r] = devices ? devices[r] : ((first < 0 ? 0 : first) + r) % ndev;
    SCGBM_REQUIRE(dv[r] >= 0 && dv[r] < ndev, "device ordinal out of range");
    for (int q = 0; q < r; ++q) SCGBM_REQUIRE(dv[q] != dv[r], "the same device is listed twice");
  }
  return dv;
}

struct RankError { int status = 0; std::string msg; };

// runs body(rank) for rank = 0..n-1, rank 0 on the calling thread; every rank gets its communicator
template <typename F>
static void run_ranks(const std::vector<int>& dev, F&& body) {
  const int n = (int)dev.size();
  uint8_t id[SCGBM_UNIQUE_ID_BYTES];
  comm_unique_id(id);
  std::vector<RankError> err(n);
  InprocGroup* group = inproc_group_create(n);
  std::vector<Comm*> comms(n, nullptr);
  auto one = [&](int r) {
    try {
      SCGBM_CUDA(cudaSetDevice(dev[r]));
      comms[r] = comm_create(id, n, r, dev[r]);
      comm_set_in_process(comms[r], group); // see comm.cu: host rendezvous before, completion after every collective
      body(r, comms[r]);
    } catch (const Error& e) { err[r].status = e.status; err[r].msg = e.what(); }
    catch (const std::bad_alloc&) { err[r].status = SCGBM_ERR_NOMEM; err[r].msg = "host allocation failed"; }
    catch (const std::exception& e) { err[r].status = SCGBM_ERR_CUDA; err[r].msg = e.what(); }
    if (err[r].status) inproc_group_abort(group);     // release the ranks waiting for this one
    // tear-down only when every rank has left its body: no collective can be in flight then, so the cross-device
    // synchronisation of ncclCommDestroy's frees cannot wait for a kernel that needs a peer
    inproc_group_finish(group);
    cudaSetDevice(dev[r]);
    cudaDeviceSynchronize();
    if (comms[r]) { comm_destroy(comms[r]); comms[r] = nullptr; }
  };
  int cur = 0;
  cudaGetDevice(&cur);
  std::vector<std::thread> th;
  for (int r = 1; r < n; ++r) th.emplace_back(one, r);
  one(0);
  for (auto& t : th) t.join();
  inproc_group_destroy(group);
  cudaSetDevice(cur);
  // report the most informative failure: a rank's own message beats "rejected on another rank"
  int pick = -1;
  for (int r = 0; r < n; ++r)
    if (err[r].status && (pick < 0 || (err[pick].msg.find("another rank") != std::string::npos &&
                                       err[r].msg.find("another rank") == std::string::npos)))
      pick = r;
  if (pick >= 0) {
    char b[32];
    snprintf(b, sizeof(b), " [rank %d of %d]", pick, n);
    throw Error(err[pick].status, err[pick].msg + b);
  }
}

void fit_multi_impl(const scgbm_fit_args* a, scgbm_fit_out* out) {
  SCGBM_REQUIRE(a && out, "NULL args");
  const int n = a->ngpu;
  SCGBM_REQUIRE(n >= 2, "fit_multi: ngpu must be >= 2");
  SCGBM_REQUIRE(a->comm == nullptr, "ngpu > 1 and comm are mutually exclusive (comm = one process per GPU)");
  SCGBM_REQUIRE(!a->y_on_device, "ngpu > 1 needs Y in host memory (each GPU uploads its own cells)");
  SCGBM_REQUIRE(a->I >= 1 && a->J >= n, "Y needs at least one cell per GPU");
  SCGBM_REQUIRE(a->Y || (a->csc_p && a->csc_i && a->csc_x), "Y is NULL (pass a dense matrix or the csc_* arrays)");
  SCGBM_REQUIRE(a->M >= 1, "M must be >= 1");
  SCGBM_REQUIRE(a->max_iter >= 1, "max.iter must be >= 1");
  SCGBM_REQUIRE(out->U && out->D && out->V && out->alpha && out->beta && out->obj, "output buffers missing");
  const std::vector<int> dev = pick_devices(n, a->devices, a->device);
  const int64_t I = a->I, J = a->J;
  const int M = a->M, nbatch = a->batch ? std::max(1, a->nbatch) : 1;

  struct RankBuf {
    std::vector<double> V, scores, Xv;                       // J_r x M: the caller's matrices have leading dimension J
    std::vector<double> U, D, loadings, alpha, obj, time, LL, Xu;   // replicated: ranks > 0 write into scratch
    std::vector<int8_t> acc;
    scgbm_fit_out o;
    int64_t j0 = 0, j1 = 0;
  };
  std::vector<RankBuf> rb(n);
  for (int r = 0; r < n; ++r) shard(J, r, n, rb[r].j0, rb[r].j1);

  run_ranks(dev, [&](int r, Comm* comm) {
    RankBuf& b = rb[r];
    const int64_t Jr = b.j1 - b.j0;
    scgbm_fit_args ar = *a;
    ar.ngpu = 0; ar.devices = nullptr;
    ar.comm = reinterpret_cast<scgbm_comm*>(comm);
    ar.device = dev[r];
    ar.J = Jr;
    if (a->Y) ar.Y = (const uint8_t*)a->Y + (size_t)b.j0 * a->ldY * dtype_size(a->y_dtype);
    else ar.csc_p = a->csc_p + b.j0;                         // absolute offsets into csc_i / csc_x stay valid
    if (a->batch) ar.batch = a->batch + b.j0;
    if (r != 0) { ar.progress = nullptr; ar.verbose = 0; }   // callbacks only on the calling thread
    scgbm_fit_out& o = b.o;
    memset(&o, 0, sizeof(o));
    b.V.assign((size_t)Jr * M, 0.0); o.V = b.V.data();
    if (out->scores) { b.scores.assign((size_t)Jr * M, 0.0); o.scores = b.scores.data(); }
    if (out->Xv) { b.Xv.assign((size_t)Jr * M, 0.0); o.Xv = b.Xv.data(); }
    o.beta = out->beta + b.j0;
    o.W = (a->return_W && out->W) ? out->W + (size_t)b.j0 * I : nullptr;
    if (r == 0) {
      o.U = out->U; o.D = out->D; o.loadings = out->loadings; o.alpha = out->alpha; o.obj = out->obj;
      o.time = out->time; o.LL = out->LL; o.accepted = out->accepted; o.Xu = out->Xu;
    } else {
      b.U.assign((size_t)I * M, 0.0); b.D.assign(M, 0.0); b.alpha.assign((size_t)I * nbatch, 0.0);
      b.obj.assign(a->max_iter, 0.0);
      o.U = b.U.data(); o.D = b.D.data(); o.alpha = b.alpha.data(); o.obj = b.obj.data();
      if (out->Xu) { b.Xu.assign((size_t)I * M, 0.0); o.Xu = b.Xu.data(); }
    }
    fit_impl(&ar, &o);
  });

  // stitch the cell-side matrices (leading dimension J in the caller's buffers)
  for (int r = 0; r < n; ++r) {
    const RankBuf& b = rb[r];
    const int64_t Jr = b.j1 - b.j0;
    for (int m = 0; m < M; ++m) {
      memcpy(out->V + (size_t)m * J + b.j0, b.V.data() + (size_t)m * Jr, Jr * sizeof(double));
      if (out->scores) memcpy(out->scores + (size_t)m * J + b.j0, b.scores.data() + (size_t)m * Jr, Jr * sizeof(double));
      if (out->Xv) memcpy(out->Xv + (size_t)m * J + b.j0, b.Xv.data() + (size_t)m * Jr, Jr * sizeof(double));
    }
  }
  const scgbm_fit_out& o0 = rb[0].o;
  out->n_obj = o0.n_obj; out->n_iter = o0.n_iter; out->hit_max_iter = o0.hit_max_iter;
  out->y_storage = o0.y_storage; out->max_Y = o0.max_Y; out->x_rank = o0.x_rank;
  out->svd_unconverged = o0.svd_unconverged; out->svd_worst_est = o0.svd_worst_est;
  out->underflow_reruns = o0.underflow_reruns; out->exact_rows = o0.exact_rows;
  out->prof = o0.prof;
  for (int r = 1; r < n; ++r) {                      // host<->device traffic of the whole call
    out->prof.h2d_bytes += rb[r].o.prof.h2d_bytes;
    out->prof.d2h_bytes += rb[r].o.prof.d2h_bytes;
    out->max_Y = std::max(out->max_Y, rb[r].o.max_Y);
  }
}

// get.se over ngpu GPUs from one caller: se_scores is per cell (local); the gene-side Gram partials are summed
// over the ranks inside scgbm_get_se (R/uncertainty.R:29-32)
void get_se_multi_impl(const scgbm_se_args* a, scgbm_se_out* out) {
  SCGBM_REQUIRE(a && out && out->se_scores && out->se_loadings, "NULL args");
  const int n = a->ngpu;
  SCGBM_REQUIRE(n >= 2 && a->comm == nullptr, "get_se_multi: ngpu >= 2 and no comm");
  SCGBM_REQUIRE(a->J >= n, "need at least one cell per GPU");
  const std::vector<int> dev = pick_devices(n, a->devices, a->device);
  const int64_t I = a->I, J = a->J;
  const int M = a->M;
  struct RankBuf { std::vector<double> scores, se_s, se_l, Xv; scgbm_se_out o; int64_t j0 = 0, j1 = 0; };
  std::vector<RankBuf> rb(n);
  for (int r = 0; r < n; ++r) shard(J, r, n, rb[r].j0, rb[r].j1);
  run_ranks(dev, [&](int r, Comm* comm) {
    RankBuf& b = rb[r];
    const int64_t Jr = b.j1 - b.j0;
    scgbm_se_args ar = *a;
    ar.ngpu = 0; ar.devices = nullptr; ar.comm = reinterpret_cast<scgbm_comm*>(comm); ar.device = dev[r];
    ar.J = Jr;
    b.scores.resize((size_t)Jr * M);
    for (int m = 0; m < M; ++m) memcpy(b.scores.data() + (size_t)m * Jr, a->scores + (size_t)m * J + b.j0, Jr * sizeof(double));
    ar.scores = b.scores.data();
    if (a->W) ar.W = a->W + (size_t)b.j0 * I;
    if (a->beta) ar.beta = a->beta + b.j0;
    if (a->batch) ar.batch = a->batch + b.j0;
    if (a->Xv) {
      b.Xv.resize((size_t)Jr * a->Mx);
      for (int m = 0; m < a->Mx; ++m) memcpy(b.Xv.data() + (size_t)m * Jr, a->Xv + (size_t)m * J + b.j0, Jr * sizeof(double));
      ar.Xv = b.Xv.data();
    }
    b.se_s.assign((size_t)Jr * M, 0.0);
    memset(&b.o, 0, sizeof(b.o));
    b.o.se_scores = b.se_s.data();
    if (r == 0) b.o.se_loadings = out->se_loadings;
    else { b.se_l.assign((size_t)I * M, 0.0); b.o.se_loadings = b.se_l.data(); }
    get_se_impl(&ar, &b.o);
  });
  out->n_singular = 0;
  for (int r = 0; r < n; ++r) {
    const RankBuf& b = rb[r];
    const int64_t Jr = b.j1 - b.j0;
    for (int m = 0; m < M; ++m) memcpy(out->se_scores + (size_t)m * J + b.j0, b.se_s.data() + (size_t)m * Jr, Jr * sizeof(double));
  }
  out->n_singular = rb[0].o.n_singular;
  out->prof = rb[0].o.prof;
}

}  // namespace scgbm
