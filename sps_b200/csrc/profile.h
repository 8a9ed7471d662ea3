// Optional per-stage CUDA-event timing of a forward (bench.py's live roofline numbers).
#pragma once
#include <cuda_runtime.h>
namespace sps {
void prof_begin(cudaStream_t st);               // first mark of a forward
void prof_mark(const char* name, cudaStream_t st);  // closes the segment `name`
bool prof_on();
}
