// Compile-only probe: registers / spills of the second streamed megakernel (nvcc -Xptxas -v).
#include "../llama-gguf_b200/csrc/stream2.cuh"
const void* s2_probe() { return (const void*)b200::stream2_decode_kernel<128, 4>; }
