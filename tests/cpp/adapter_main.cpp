// adapter_main.cpp -- exercises the header-only C++ adapter (rsfusion::TSDFSystem / TSDFGrid / SE3)
// the way main/offline_eval.cc drives the reference: queue frames, render, download.  The Python
// test feeds it a frame file and compares its dumps with the ctypes path (which is parity-checked
// against the oracle).
//
// usage: adapter_main frames.bin out_prefix
// frames.bin: int32 n,H,W | float K[4] | float voxel,trunc,max_depth | n x { float pose[7] (qw qx qy qz tx ty tz),
//             u8 rgb[H*W*3], float depth[H*W], float ht[H*W], float lt[H*W] } | float mats[2][16] (row-major 4x4)
#define RSFUSION_GLOBAL_NAMES
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "rsfusion/tsdf_system.hpp"

template <class T>
static void dump(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}

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
  try {
    const CameraIntrinsics<float> intr(K[0], K[1], K[2], K[3]);
    TSDFSystem sys(par[0], par[1], par[2], intr);
    std::vector<uint8_t> rgb(P * 3);
    std::vector<float> depth(P), ht(P), lt(P);
    SE3<float> last;
    for (int i = 0; i < n; ++i) {
      float pose[7];
      in.read((char*)pose, sizeof pose);
      in.read((char*)rgb.data(), (std::streamsize)rgb.size());
      in.read((char*)depth.data(), (std::streamsize)(P * 4));
      in.read((char*)ht.data(), (std::streamsize)(P * 4));
      in.read((char*)lt.data(), (std::streamsize)(P * 4));
      last = SE3<float>({pose[0], pose[1], pose[2], pose[3]}, {pose[4], pose[5], pose[6]});
      // the buffers are reused right after the call returns, exactly like offline_eval's loop
      sys.Integrate(last, rsfusion::Image(rgb.data(), H, W, rsfusion::kType8UC3), rsfusion::Image(depth.data(), H, W, rsfusion::kType32FC1),
                    rsfusion::Image(ht.data(), H, W, rsfusion::kType32FC1), rsfusion::Image(lt.data(), H, W, rsfusion::kType32FC1));
    }
    sys.Flush();
    std::vector<uint8_t> rgba(P * 4), normal(P * 4);
    std::vector<float> rdepth(P);
    sys.Render(CameraParams(intr, H, W), last, rgba.data(), normal.data(), par[2] * 2, rdepth.data());
    dump(out + ".rgba", rgba);
    dump(out + ".normal", normal);
    dump(out + ".depth", rdepth);
    sys.DownloadAll(out + ".tsdf");
    const auto q = sys.Query(BoundingCube<float>{0.f, 2.5f, 0.f, 4.f, 0.f, 1.5f});
    dump(out + ".query", q);
    // SE3 from 4x4 matrices (Eigen's Quaternion(Matrix3) algorithm) and composition
    float mats[2][16];
    in.read((char*)mats, sizeof mats);
    std::vector<float> se3;
    const SE3<float> A(mats[0]), B(mats[1]);
    for (const SE3<float>& T : {A, B, A * B, A.Inverse()}) {
      const auto p = T.as_pose();
      se3.insert(se3.end(), p.begin(), p.end());
    }
    dump(out + ".se3", se3);
    std::printf("active_blocks %d rows %zu\n", sys.grid().NumActiveBlock(), q.size());
    sys.terminate();
    // error behaviour: wrong type is rejected with an exception, not an assert
    bool threw = false;
    try {
      TSDFGrid g(par[0], par[1]);
      g.Integrate(rsfusion::Image(depth.data(), H, W, rsfusion::kType32FC1), rsfusion::Image(depth.data(), H, W, rsfusion::kType32FC1),
                  rsfusion::Image(ht.data(), H, W, rsfusion::kType32FC1), rsfusion::Image(lt.data(), H, W, rsfusion::kType32FC1), 4.f, intr, last);
    } catch (const rsfusion::Error& e) {
      threw = e.code == RSF_ERR_INVALID;
    }
    std::printf("type_check %s\n", threw ? "ok" : "MISSING");
  } catch (const std::exception& e) {
    std::fprintf(stderr, "adapter_main: %s\n", e.what());
    return 1;
  }
  return 0;
}
