/*
 * miniomm_backend.h -- TEST INFRASTRUCTURE ONLY: what mini_openmm.cpp (plain host C++) needs from plugin_kernels.cpp
 * (the TU that holds the reference kernels and, in the CUDA flavour, everything that touches the device).
 */
#ifndef MINIOMM_BACKEND_H_
#define MINIOMM_BACKEND_H_
#include <cstdint>
#include <string>

namespace miniomm {
struct KernelEntry;
void setDefine(const std::string &name, long value, void *stream);                  /* NUM_ATOMS ... -> variables */
KernelEntry *findKernel(const std::string &module, const std::string &name, int numTG);   /* nullptr if unknown */
void launchKernel(KernelEntry *k, void **args, int grid, int block, unsigned sharedSize, void *stream);
void *deviceAlloc(size_t bytes);                                                     /* zero-initialised */
void deviceFree(void *p);
void copyToDevice(void *dst, const void *src, size_t bytes, void *stream);
void copyToHost(void *dst, const void *src, size_t bytes, void *stream);             /* synchronises the stream */
int numThreadBlocks();                                                               /* 4 x SMs [OMM-mem]; 4 x 148 on the host */
bool isCuda();
/* the constraint stand-in (oracle/constraint_standin.h) on this flavour's memory */
struct Standin { int numClusters = 0, iterations = 0; int32_t *offset = nullptr, *atoms = nullptr; double *distance = nullptr; };
void standinPositions(const Standin &s, void *posq, void *corr, void *velm, void *posDelta, void *stream);
void standinVelocities(const Standin &s, void *posq, void *corr, void *velm, void *stream);
}   // namespace miniomm

#endif
