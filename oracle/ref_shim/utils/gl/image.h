// utils/gl/image.h -- stand-in for the reference's OpenGL texture bridge (test infrastructure).
// TSDFGrid::RayCast (reference utils/tsdf/voxel_tsdf.cu:885-902) only calls LoadCuda on the optional
// outputs; GLEW / GL are absent from this image, and the harness reads img_tsdf_rgba_ directly.
#ifndef RSF_REF_SHIM_GL_IMAGE_
#define RSF_REF_SHIM_GL_IMAGE_
#include <cuda_runtime.h>
class GLImage8UC4 {
 public:
  void LoadCuda(const void*, cudaStream_t = nullptr) {}
};
class GLImage32FC1 {
 public:
  void LoadCuda(const void*, cudaStream_t = nullptr) {}
};
#endif
