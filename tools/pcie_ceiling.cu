// pcie_ceiling.cu -- torch-free host<->device copy ceiling with the access pattern of b200_adam_host_step_f32.
//
//   build/pcie_ceiling [--mib-per-array 512] [--slice-mib 16] [--reps 3] [--flags 0|portable|wc|portable+wc] [--device D]
//                      [--start-at UNIX_SECONDS]
//
// Moves 6 input arrays host->device and 6 output arrays device->host in slices on two streams (one cudaMemcpyAsync per
// slice per array, nothing else: no kernels, no dependencies), exactly the copy schedule of the e2e leg of bench.py.
// Prints one JSON line: GB/s per direction.  Run N copies at once (one per GPU, e.g. from tools/pcie_ceiling_all.sh, all
// given the same --start-at) to measure what the HOST gives N GPUs copying simultaneously.  The pinned allocation flags
// are selectable to test whether portable / write-combined memory moves the ceiling (write-combined applies to the
// INPUT arrays only: the CPU must be able to read the outputs at cached speed).
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

int main(int argc, char **argv) {
  size_t mib = 512, slice_mib = 16;
  int reps = 3, device = 0;
  double start_at = 0;
  unsigned in_flags = cudaHostAllocDefault, out_flags = cudaHostAllocDefault;
  const char *flag_name = "default";
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--mib-per-array")) mib = strtoull(argv[i + 1], nullptr, 10);
    else if (!strcmp(argv[i], "--slice-mib")) slice_mib = strtoull(argv[i + 1], nullptr, 10);
    else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--start-at")) start_at = atof(argv[i + 1]);
    else if (!strcmp(argv[i], "--flags")) {
      flag_name = argv[i + 1];
      if (strstr(flag_name, "portable")) in_flags |= cudaHostAllocPortable, out_flags |= cudaHostAllocPortable;
      if (strstr(flag_name, "wc")) in_flags |= cudaHostAllocWriteCombined;
    }
  }
  CK(cudaSetDevice(device));
  const size_t bytes = mib << 20, slice = slice_mib << 20;
  std::vector<char *> h_in(6), h_out(6), d_in(6), d_out(6);
  for (int a = 0; a < 6; ++a) {
    CK(cudaHostAlloc((void **)&h_in[a], bytes, in_flags));
    CK(cudaHostAlloc((void **)&h_out[a], bytes, out_flags));
    memset(h_in[a], 1, bytes);   // first touch on the allocating thread (NUMA placement follows the CPU affinity)
    memset(h_out[a], 0, bytes);
    CK(cudaMalloc((void **)&d_in[a], slice * 3));   // 3 staging slots per array, like the pipeline
    CK(cudaMalloc((void **)&d_out[a], slice * 3));
  }
  cudaStream_t s_in, s_out;
  CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  if (start_at > 0) {   // crude cross-process barrier: everybody starts copying at the same wall-clock second
    using namespace std::chrono;
    const double now = duration<double>(system_clock::now().time_since_epoch()).count();
    if (start_at > now) std::this_thread::sleep_for(duration<double>(start_at - now));
  }
  double best = 1e30, sum = 0;
  for (int r = 0; r < reps + 1; ++r) {
    const auto w0 = std::chrono::steady_clock::now();
    size_t k = 0;
    for (size_t off = 0; off < bytes; off += slice, ++k) {
      const size_t len = bytes - off < slice ? bytes - off : slice;
      for (int a = 0; a < 6; ++a) CK(cudaMemcpyAsync(d_in[a] + (k % 3) * slice, h_in[a] + off, len, cudaMemcpyHostToDevice, s_in));
      for (int a = 0; a < 6; ++a) CK(cudaMemcpyAsync(h_out[a] + off, d_out[a] + (k % 3) * slice, len, cudaMemcpyDeviceToHost, s_out));
    }
    CK(cudaStreamSynchronize(s_in));
    CK(cudaStreamSynchronize(s_out));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    if (r == 0) continue;   // warm-up pass
    best = sec < best ? sec : best;
    sum += sec;
  }
  const double gb = 6.0 * bytes / 1e9;
  printf("{\"device\": %d, \"flags\": \"%s\", \"mib_per_array\": %zu, \"slice_mib\": %zu, \"GBps_each_direction_best\": %.2f, "
         "\"GBps_each_direction_mean\": %.2f, \"gelem_per_s_at_24B_each_way\": %.3f}\n",
         device, flag_name, mib, slice_mib, gb / best, gb / (sum / reps), gb / (sum / reps) / 24.0);
  return 0;
}
