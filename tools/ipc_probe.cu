// Probe: does CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) work between two processes on this
// box, and can a kernel in the opener store into the exporter's buffer (same GPU, or peer GPU over NVLink
// when >= 2 devices are visible)?  Decides whether the shard exchange can use peer stores.
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("FAIL %s: %s\n", #x, cudaGetErrorString(e));                          \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

__global__ void fill(unsigned* p, unsigned n, unsigned v) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v + i;
  __threadfence_system();
}

int main() {
  int p2c[2], c2p[2];
  if (pipe(p2c) || pipe(c2p)) return 2;
  pid_t pid = fork();  // before any CUDA call
  const unsigned n = 1 << 20;
  if (pid == 0) {
    cudaIpcMemHandle_t h;
    if (read(p2c[0], &h, sizeof(h)) != (ssize_t)sizeof(h)) return 3;
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    CK(cudaSetDevice(nd > 1 ? 1 : 0));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    fill<<<(n + 255) / 256, 256>>>((unsigned*)p, n, 7u);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaIpcCloseMemHandle(p));
    char ok = 1;
    if (write(c2p[1], &ok, 1) != 1) return 4;
    return 0;
  }
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  CK(cudaSetDevice(0));
  unsigned* d = nullptr;
  CK(cudaMalloc(&d, n * 4));
  CK(cudaMemset(d, 0, n * 4));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, d));
  if (write(p2c[1], &h, sizeof(h)) != (ssize_t)sizeof(h)) return 5;
  char ok = 0;
  ssize_t r = read(c2p[0], &ok, 1);
  int st = 0;
  waitpid(pid, &st, 0);
  unsigned probe[2] = {0, 0};
  CK(cudaMemcpy(&probe[0], d, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&probe[1], d + n - 1, 4, cudaMemcpyDeviceToHost));
  bool good = r == 1 && ok == 1 && probe[0] == 7u && probe[1] == 7u + n - 1;
  printf("ipc_probe devices=%d opener_device=%d result=%s (child status %d, first=%u last=%u)\n", nd, nd > 1 ? 1 : 0,
         good ? "OK" : "BAD", st, probe[0], probe[1]);
  return good ? 0 : 1;
}
