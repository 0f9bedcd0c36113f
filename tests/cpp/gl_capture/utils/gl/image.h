// Test stand-in for the reference's utils/gl/image.h (no OpenGL in this image): GLImage8UC4 whose LoadCuda copies the
// CUDA image into a host vector instead of a GL texture, so the test can look at what TSDFSystem::Render produced.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>
class GLImage8UC4 {
 public:
  void BindImage(int height, int width, const void* = nullptr) {
    height_ = height;
    width_ = width;
    host.assign((size_t)height * width * 4, 0);
  }
  void LoadCuda(const void* data, cudaStream_t stream = nullptr) {
    cudaMemcpyAsync(host.data(), data, host.size(), cudaMemcpyDeviceToHost, stream);
    ++loads;
  }
  int GetHeight() { return height_; }
  int GetWidth() { return width_; }
  std::vector<uint8_t> host;
  int loads = 0;

 private:
  int height_ = 0, width_ = 0;
};
