// host-only stand-ins for the few CUDA runtime calls arrow_bridge.cpp makes (TEST INFRASTRUCTURE)
// host-only debug stubs
#pragma once
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
typedef int cudaError_t; typedef void* cudaStream_t;
#define cudaSuccess 0
#define cudaHostAllocPortable 1
#define cudaHostAllocWriteCombined 4
#define cudaStreamNonBlocking 1
#define cudaMemcpyHostToDevice 1
#define cudaMemcpyDeviceToHost 2
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n); return *p ? 0 : 1; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "stub"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
// PB_STUB_STREAMS_OK: stream creation succeeds (the unary-sweep glue runs end to end against the CPU doubles of harness_tail.inc);
// otherwise it fails, which keeps the binary operations from going anywhere near their (failing) device stubs
extern int pb_stub_streams_ok;
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 0; return pb_stub_streams_ok ? 0 : 1; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { *p = malloc(n); return 0; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return 0; }
// "device" memory is malloc'ed host memory here (dev_alloc in harness_tail.inc), so a copy is a copy
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { if (n) memcpy(d, s, n); return 0; }
typedef void* cudaEvent_t;
#define cudaEventDisableTiming 2
#define cudaHostRegisterPortable 1
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = 0; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return 0; }
static inline cudaError_t cudaHostUnregister(void*) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaMemsetAsync(void*, int, size_t, cudaStream_t) { return 0; }
