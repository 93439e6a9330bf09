/* Stub for boxes whose g++ has no OpenMP runtime files (libgomp.spec / omp.h): bench.py's torch_compile_gpu leg points
 * torch-inductor's host compiler at tools/inductor_cxx/g++-noomp, which drops -fopenmp / -lgomp and finds this header.
 * Only the two functions torch/csrc/inductor/cpp_prefix.h references.  Benchmark plumbing for the REFERENCE arm — the
 * 0-d host-scalar arithmetic of the reference's step is the only thing inductor compiles for the CPU. */
#ifndef CONSOLVER_BENCH_OMP_STUB_H_
#define CONSOLVER_BENCH_OMP_STUB_H_
static inline int omp_get_thread_num(void) { return 0; }
static inline int omp_get_max_threads(void) { return 1; }
static inline int omp_get_num_threads(void) { return 1; }
static inline int omp_in_parallel(void) { return 0; }
#endif
