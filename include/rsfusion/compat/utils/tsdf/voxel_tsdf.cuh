// Include overlay: put  -I<rsfusion>/include/rsfusion/compat -I<rsfusion>/include  BEFORE the RA-SLAM source root and the
// reference's own  #include "utils/tsdf/voxel_tsdf.cuh"  (modules/tsdf_module.h:15, examples/tsdf/offline.cc, ...)
// resolves to this file: the same class and method names (TSDFGrid, SE3, CameraIntrinsics, CameraParams, BoundingCube,
// VoxelSpatialTSDF, VoxelSpatialTSDFSEGM) backed by librsfusion.so instead of utils/tsdf/*.cu.  modules/tsdf_module.cc
// compiles unmodified against it (tests/test_reference_sources_compile.py does exactly that).
// The caller's tree supplies OpenCV (cv::Mat), Eigen and utils/gl/image.h (GLImage8UC4), as it does for the reference.
#pragma once
#include <opencv2/opencv.hpp>

#include <Eigen/Dense>
#include <Eigen/Geometry>

#include "utils/gl/image.h"   // GLImage8UC4 (the reference's OpenGL bridge; voxel_tsdf.cuh:10)

#ifndef RSFUSION_WITH_OPENCV
#define RSFUSION_WITH_OPENCV
#endif
#ifndef RSFUSION_WITH_EIGEN
#define RSFUSION_WITH_EIGEN
#endif
#ifndef RSFUSION_WITH_GL
#define RSFUSION_WITH_GL
#endif
#ifndef RSFUSION_GLOBAL_NAMES
#define RSFUSION_GLOBAL_NAMES
#endif
#include "rsfusion/tsdf_grid.hpp"
