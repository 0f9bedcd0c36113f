// ref_module_main.cpp -- drives the REFERENCE's own TSDFSystem (modules/tsdf_module.{h,cc}, compiled UNMODIFIED from
// /root/reference) on top of this library: its  #include "utils/tsdf/voxel_tsdf.cuh"  resolves to the include overlay
// include/rsfusion/compat/, i.e. to rsfusion::TSDFGrid.  Proves the drop-in claim for the calls tsdf_module.cc makes:
// Integrate(cv::Mat...) :107, RayCast(..., GLImage8UC4*, GLImage8UC4*) :48,54, GatherVoxels :41, GatherValidSemantic :59,
// GatherValidMesh(std::vector<Eigen::Vector3f>*, ...) :72.
//
// usage: ref_module_main frames.bin out_prefix      (frames.bin as for adapter_main)
#include <chrono>
#include <cstdio>
#include <fstream>
#include <thread>
#include <vector>

#include "modules/tsdf_module.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  const std::string out = argv[2];
  int32_t hdr[3];
  float K[4], par[3];
  in.read((char*)hdr, sizeof hdr);
  in.read((char*)K, sizeof K);
  in.read((char*)par, sizeof par);
  const int n = hdr[0], H = hdr[1], W = hdr[2];
  const size_t P = (size_t)H * W;
  const CameraIntrinsics<float> intr(K[0], K[1], K[2], K[3]);
  TSDFSystem sys(par[0], par[1], par[2], intr);
  std::vector<uint8_t> rgb(P * 3);
  std::vector<float> depth(P), ht(P), lt(P);
  SE3<float> last = SE3<float>::Identity();
  for (int i = 0; i < n; ++i) {
    float pose[7];
    in.read((char*)pose, sizeof pose);
    in.read((char*)rgb.data(), (std::streamsize)rgb.size());
    in.read((char*)depth.data(), (std::streamsize)(P * 4));
    in.read((char*)ht.data(), (std::streamsize)(P * 4));
    in.read((char*)lt.data(), (std::streamsize)(P * 4));
    last = SE3<float>(Eigen::Quaternionf(pose[0], pose[1], pose[2], pose[3]), Eigen::Vector3f(pose[4], pose[5], pose[6]));
    sys.Integrate(last, cv::Mat(H, W, CV_8UC3, rgb.data()), cv::Mat(H, W, CV_32FC1, depth.data()), cv::Mat(H, W, CV_32FC1, ht.data()),
                  cv::Mat(H, W, CV_32FC1, lt.data()));
  }
  // the reference's queue has no flush: Render / DownloadAll take mtx_read_ whenever the worker lets go of it, so give the
  // worker time to drain (its Run() loop spins on the queue)
  for (int spin = 0; spin < 200; ++spin) std::this_thread::sleep_for(std::chrono::milliseconds(10));
  GLImage8UC4 rgba, normal;
  rgba.BindImage(H, W);
  normal.BindImage(H, W);
  sys.Render(CameraParams(intr, H, W), last, &rgba, &normal);
  std::ofstream(out + ".rgba", std::ios::binary).write((const char*)rgba.host.data(), (std::streamsize)rgba.host.size());
  std::ofstream(out + ".normal", std::ios::binary).write((const char*)normal.host.data(), (std::streamsize)normal.host.size());
  sys.DownloadAll(out + ".tsdf");
  sys.DownloadAllMesh(out + ".vertices", out + ".indices", out + ".vprob");
  const auto q = sys.Query(BoundingCube<float>{0.f, 2.5f, 0.f, 4.f, 0.f, 1.5f});
  std::ofstream(out + ".query", std::ios::binary).write((const char*)q.data(), (std::streamsize)(q.size() * sizeof(VoxelSpatialTSDF)));
  std::printf("query_rows %zu loads %d %d\n", q.size(), rgba.loads, normal.loads);
  // (no sys.terminate() here: the reference's destructor calls it, modules/tsdf_module.cc:20, and its terminate() joins
  // the worker unconditionally, :127-133 -- a second call throws std::system_error)
  return 0;
}
