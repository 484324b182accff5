// Multi-GPU entry points of the C ABI: what a non-Python host needs for the sharded step.
// The reference has no distributed code at all; these replace nothing in it - they are the two exchanges the sharded
// path adds around src/multigrid_model.py:301-324 (halo rows before / after the operator products, sums of partials
// and gradients).  NCCL is NOT linked: its entry points are resolved at run time from the library the process
// already uses (in a torch process: the libnccl.so.2 torch ships, which owns the communicator the caller passes),
// so the library has no link-time dependency and never mixes two NCCL instances.
#include <dlfcn.h>
#include <nccl.h>          // types and enums only

#include <mutex>

#include "ep_common.cuh"
#include "../../include/eigenpinns_b200.h"

namespace {

struct NcclApi {
  decltype(&ncclSend) send = nullptr;
  decltype(&ncclRecv) recv = nullptr;
  decltype(&ncclAllReduce) all_reduce = nullptr;
  decltype(&ncclGroupStart) group_start = nullptr;
  decltype(&ncclGroupEnd) group_end = nullptr;
  decltype(&ncclGetErrorString) error_string = nullptr;
  decltype(&ncclGetVersion) get_version = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (!dlsym(RTLD_DEFAULT, "ncclSend")) {
      // by SONAME: returns the instance that is already mapped (torch loads its own with RTLD_LOCAL), else the system one
      h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) return;
    }
    auto sym = [&](const char* name) { return h ? dlsym(h, name) : dlsym(RTLD_DEFAULT, name); };
    a.send = reinterpret_cast<decltype(a.send)>(sym("ncclSend"));
    a.recv = reinterpret_cast<decltype(a.recv)>(sym("ncclRecv"));
    a.all_reduce = reinterpret_cast<decltype(a.all_reduce)>(sym("ncclAllReduce"));
    a.group_start = reinterpret_cast<decltype(a.group_start)>(sym("ncclGroupStart"));
    a.group_end = reinterpret_cast<decltype(a.group_end)>(sym("ncclGroupEnd"));
    a.error_string = reinterpret_cast<decltype(a.error_string)>(sym("ncclGetErrorString"));
    a.get_version = reinterpret_cast<decltype(a.get_version)>(sym("ncclGetVersion"));
    a.ok = a.send && a.recv && a.all_reduce && a.group_start && a.group_end;
  });
  return a;
}

int fail(const char* where, ncclResult_t r) {
  char msg[256];
  snprintf(msg, sizeof(msg), "%s: NCCL error %d (%s)", where, (int)r, api().error_string ? api().error_string(r) : "?");
  ep::set_error("%s", msg);
  return EP_ERR_CUDA;
}

}  // namespace

extern "C" {

int ep_dist_nccl_version(void) {
  int v = 0;
  if (!api().ok || !api().get_version || api().get_version(&v) != ncclSuccess) return 0;
  return v;
}

int ep_halo_exchange_f32(void* nccl_comm, int n_peers, const int* peer_rank, const int* send_offset, const int32_t* send_idx,
                         const int* recv_offset, int k, const float* rows, int ld, float* send_buf, float* halo_rows,
                         ep_stream_t stream) {
  EP_REQUIRE(n_peers >= 0 && k > 0 && ld >= k, "bad size");
  if (n_peers == 0) return EP_OK;
  EP_REQUIRE(nccl_comm && peer_rank && send_offset && recv_offset && rows && halo_rows, "null pointer");
  EP_REQUIRE(ld == k, "halo rows are received in place: the row block must be dense (ld == k)");
  if (!api().ok) { ep::set_error("ep_halo_exchange_f32: no NCCL library in this process (libnccl.so.2 not found)"); return EP_ERR_UNSUPPORTED; }
  const int n_send = send_offset[n_peers];
  if (n_send > 0) {
    EP_REQUIRE(send_idx && send_buf, "null pointer");
    const int rc = ep_gather_rows_f32(n_send, k, send_idx, rows, ld, send_buf, k, stream);     // pack the boundary rows
    if (rc != EP_OK) return rc;
  }
  cudaStream_t st = ep::as_stream(stream);
  ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
  ncclResult_t r = api().group_start();
  if (r != ncclSuccess) return fail("ncclGroupStart", r);
  for (int p = 0; p < n_peers && r == ncclSuccess; ++p) {
    const size_t ns = (size_t)(send_offset[p + 1] - send_offset[p]) * k, nr = (size_t)(recv_offset[p + 1] - recv_offset[p]) * k;
    if (ns) r = api().send(send_buf + (size_t)send_offset[p] * k, ns, ncclFloat, peer_rank[p], comm, st);
    if (nr && r == ncclSuccess) r = api().recv(halo_rows + (size_t)recv_offset[p] * k, nr, ncclFloat, peer_rank[p], comm, st);
  }
  const ncclResult_t e = api().group_end();
  if (r != ncclSuccess) return fail("ncclSend/ncclRecv", r);
  if (e != ncclSuccess) return fail("ncclGroupEnd", e);
  return EP_OK;
}

int ep_allreduce_sum_f64(void* nccl_comm, size_t count, double* buf, ep_stream_t stream) {
  EP_REQUIRE(nccl_comm && (buf || count == 0), "null pointer");
  if (count == 0) return EP_OK;
  if (!api().ok) { ep::set_error("ep_allreduce_sum_f64: no NCCL library in this process"); return EP_ERR_UNSUPPORTED; }
  const ncclResult_t r = api().all_reduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(nccl_comm), ep::as_stream(stream));
  return r == ncclSuccess ? EP_OK : fail("ncclAllReduce", r);
}

int ep_allreduce_sum_f32(void* nccl_comm, size_t count, float* buf, ep_stream_t stream) {
  EP_REQUIRE(nccl_comm && (buf || count == 0), "null pointer");
  if (count == 0) return EP_OK;
  if (!api().ok) { ep::set_error("ep_allreduce_sum_f32: no NCCL library in this process"); return EP_ERR_UNSUPPORTED; }
  const ncclResult_t r = api().all_reduce(buf, buf, count, ncclFloat, ncclSum, static_cast<ncclComm_t>(nccl_comm), ep::as_stream(stream));
  return r == ncclSuccess ? EP_OK : fail("ncclAllReduce", r);
}

}  // extern "C"
