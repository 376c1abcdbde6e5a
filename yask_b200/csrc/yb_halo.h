// Multi-GPU halo exchange over NVLink peer memory (one process per GPU, CUDA IPC).
// Replaces StencilContext::exchange_halos (/root/reference/src/kernel/lib/halo.cpp:80-491).
#pragma once
#include "yb_core.h"

namespace yb {

int halo_prepare(Solution& s);                       // allocate sync flags, neighbour table
void halo_mark_dirty(Solution& s, int var);          // host wrote into a var (dirty protocol, yk_var.cpp:122-152)
int halo_exchange_all(Solution& s, cudaStream_t st); // exchange every dirty var/step, complete in stream order
int halo_run_stage(Solution& s, int stage, int64_t t, cudaStream_t st);  // wait(previous) -> exterior -> push -> interior
int halo_finish(Solution& s, cudaStream_t st);       // enqueue the wait of the exchange still in flight (end of run_solution)

}  // namespace yb
