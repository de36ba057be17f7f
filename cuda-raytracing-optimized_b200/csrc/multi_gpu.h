// multi_gpu.h -- sample-sharded rendering on several GPUs of one box INSIDE the library (the reference is single-GPU,
// kernels.cu:652-664): setRendererGpus(N) makes the next initRenderer bring up N-1 worker host threads, one per further device.
// Every device holds the whole scene (uploaded and indexed by its own thread from the caller's host arrays) and renders
// ns/N samples of EVERY pixel on its own RNG stream (stream g = wang_hash(pixel + g*nx*ny), the reference's stream being g = 0);
// runRenderer ends with ONE ncclReduce of the un-normalised float4 sums to device 0 and fb = sum / ns there. Nothing is exchanged
// per bounce. NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it and a single-GPU
// caller never touches it.
#pragma once

struct kernel_scene;
struct camera;

bool crtMultiGpuStart(const kernel_scene& sc, const camera& cam, int nx, int ny, int maxDepth, int gpus, int mainDevice, unsigned int streamBase);
// Runs the workers' shares beside the caller's own (`runOwn`), reduces, finalizes into the caller's frame buffer.
void crtMultiGpuRun(int nsTotal, void (*runOwn)(int ns));
void crtMultiGpuStop();
int crtMultiGpuCount(); // devices in use by this thread's renderer (1 = single)
