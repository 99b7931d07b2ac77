#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c, cudaGraphConditionalHandle h) { int v = ++(*c); if (v >= 5) cudaGraphSetConditional(h, 0); }
int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  int* c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t n; cudaError_t e = cudaGraphAddNode(&n, g, nullptr, 0, &p); printf("add %d\n", e);
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  e = cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed); printf("begin %d\n", e);
  body<<<1,1,0,s>>>(c, h);
  e = cudaStreamEndCapture(s, nullptr); printf("end %d\n", e);
  cudaGraphExec_t x; e = cudaGraphInstantiate(&x, g, 0); printf("inst %d\n", e);
  e = cudaGraphLaunch(x, s); cudaStreamSynchronize(s);
  int hc; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost); printf("count %d (%s)\n", hc, cudaGetErrorString(cudaGetLastError()));
}
