// TEST INFRASTRUCTURE ONLY -- see cpu_emu.h.
#include "cpu_emu.h"

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;
pthread_barrier_t emu_block_barrier;
pthread_barrier_t emu_warp_barrier[64];
volatile uint32_t emu_shfl_buf[64][32];
unsigned char* emu_dyn_smem = nullptr;

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads % 32 != 0 || nthreads > 1024) {
    fprintf(stderr, "emu: block size %u unsupported\n", nthreads);
    abort();
  }
  std::vector<unsigned char> dyn(smem + 64);
  emu_dyn_smem = dyn.data() + (64 - (reinterpret_cast<uintptr_t>(dyn.data()) & 63)) % 64;
  pthread_barrier_init(&emu_block_barrier, nullptr, nthreads);
  for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_init(&emu_warp_barrier[w], nullptr, 32);
  // persistent workers: one host thread per CUDA thread of a block, looping over the blocks
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t) {
    pool.emplace_back([&, t]() {
      blockDim = block;
      gridDim = grid;
      threadIdx.x = t % block.x;
      threadIdx.y = (t / block.x) % block.y;
      threadIdx.z = t / (block.x * block.y);
      for (unsigned long long b = 0; b < nblocks; ++b) {
        blockIdx.x = b % grid.x;
        blockIdx.y = (b / grid.x) % grid.y;
        blockIdx.z = b / ((unsigned long long)grid.x * grid.y);
        body();
        pthread_barrier_wait(&emu_block_barrier);  // block boundary: statics are reused
      }
    });
  }
  for (auto& th : pool) th.join();
  pthread_barrier_destroy(&emu_block_barrier);
  for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_destroy(&emu_warp_barrier[w]);
  emu_dyn_smem = nullptr;
}
