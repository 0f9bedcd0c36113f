// Include overlay for the reference's utils/cuda/camera.cuh: the types live in rsfusion/tsdf_grid.hpp (see
// utils/tsdf/voxel_tsdf.cuh in this directory).
#pragma once
#include "utils/tsdf/voxel_tsdf.cuh"
