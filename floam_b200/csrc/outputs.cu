// On-disk outputs of the odometry node: what the reference writes when it exits —
//   SaveMerged / SavePosesHomogeneousBALM   src/odomEstimationNode.cpp:66-121
//   SavePosegraph / SaveOdom                src/utils.cpp:3-106
// The text formats are iostream formats (default precision 6, Eigen's aligned matrix printing, Boost.Format directives that only set
// stream state), reproduced byte for byte; the PCD files are PCL 1.8's binary PointXYZI layout (x y z intensity, 16 bytes per point).
// SaveMerged's transform + merge + VoxelGrid run on the device; everything else is host I/O and needs no context.
#include <sys/stat.h>
#include <sys/types.h>

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "context.cuh"
#include "odom_math.cuh"

using namespace floam;

namespace {

bool make_dirs(const std::string& dir) {   // boost::filesystem::create_directories
  for (size_t i = 1; i <= dir.size(); ++i)
    if (i == dir.size() || dir[i] == '/') ::mkdir(dir.substr(0, i).c_str(), 0777);
  struct stat st;
  return ::stat(dir.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

// pcl::io::savePCDFileBinary<pcl::PointXYZI> (PCL 1.8.1 PCDWriter::writeBinary): header of generateHeader + "DATA binary", then the
// registered fields of every point back to back (the padding of the 32-byte struct is not written)
int write_pcd(const std::string& path, const float* xyzi16, size_t n) {
  std::FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return FLOAM_ERR_ARG;
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n");
  std::fprintf(f, "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", n, n);
  const bool ok = n == 0 || std::fwrite(xyzi16, 16, n, f) == n;
  std::fclose(f);
  return ok ? FLOAM_OK : FLOAM_ERR_ARG;
}
int write_pcd_xyzi(const std::string& path, const floam_point_xyzi* pts, size_t n) {
  std::vector<float> rec(4 * n);
  for (size_t i = 0; i < n; ++i) { rec[4 * i] = pts[i].x; rec[4 * i + 1] = pts[i].y; rec[4 * i + 2] = pts[i].z; rec[4 * i + 3] = pts[i].intensity; }
  return write_pcd(path, rec.data(), n);
}

// ros::Time(double) (TimeBase::fromSec) and the pcl stamp round trip of pcl_conversions (microseconds)
void ros_time(double t, unsigned int* sec, unsigned int* nsec) {
  const long long sec64 = (long long)std::floor(t);
  unsigned int s = (unsigned int)sec64;
  unsigned int ns = (unsigned int)std::llround((t - (double)s) * 1e9);
  s += ns / 1000000000u;
  ns %= 1000000000u;
  *sec = s; *nsec = ns;
}

// operator<<(ostream, Matrix4d) with Eigen's default IOFormat: every coefficient right-aligned to the widest one
void print_matrix4(std::ostream& s, const double* M) {   // row-major
  std::streamsize width = 0;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      std::stringstream sstr;
      sstr.copyfmt(s);
      sstr << M[i * 4 + j];
      width = std::max<std::streamsize>(width, (std::streamsize)sstr.str().length());
    }
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      if (j) s << " ";
      if (width) s.width(width);
      s << M[i * 4 + j];
    }
    if (i < 3) s << "\n";
  }
}

void rotation_of(const double* M, double* R9) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R9[r * 3 + c] = M[r * 4 + c]; }

// Eigen::Affine3d::inverse() * other (general 3x3 inverse by cofactors, then the affine product)
void affine_between(const double* A, const double* B, double* O) {
  double L[9], Li[9];
  rotation_of(A, L);
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return L[i1 * 3 + j1] * L[i2 * 3 + j2] - L[i1 * 3 + j2] * L[i2 * 3 + j1];
  };
  const double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const double invdet = 1.0 / (c0 * L[0] + c1 * L[3] + c2 * L[6]);
  Li[0] = c0 * invdet; Li[1] = c1 * invdet; Li[2] = c2 * invdet;
  Li[3] = cof(0, 1) * invdet; Li[4] = cof(1, 1) * invdet; Li[5] = cof(2, 1) * invdet;
  Li[6] = cof(0, 2) * invdet; Li[7] = cof(1, 2) * invdet; Li[8] = cof(2, 2) * invdet;
  double ti[3];
  for (int r = 0; r < 3; ++r) ti[r] = -(Li[r * 3] * A[3] + Li[r * 3 + 1] * A[7] + Li[r * 3 + 2] * A[11]);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) O[r * 4 + c] = Li[r * 3] * B[c] + Li[r * 3 + 1] * B[4 + c] + Li[r * 3 + 2] * B[8 + c];
    O[r * 4 + 3] = Li[r * 3] * B[3] + Li[r * 3 + 1] * B[7] + Li[r * 3 + 2] * B[11] + ti[r];
  }
  O[12] = O[13] = O[14] = 0.0; O[15] = 1.0;
}

bool args_ok(const char* dir, const double* poses, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets, int n) {
  if (!dir || n < 0 || (n > 0 && (!poses || !stamps || !offsets))) return false;
  for (int i = 0; i < n; ++i)
    if (offsets[i + 1] < offsets[i] || (offsets[i + 1] > offsets[i] && !clouds)) return false;
  return true;
}

// pcl::transformPointCloud(cloud, out, Eigen::Affine3d): double arithmetic per coefficient, float store (PCL 1.8.1 transforms.hpp)
__global__ void transform_affine_kernel(const PointI* __restrict__ in, int n, const double* __restrict__ M, P4* __restrict__ out) {
  pdl_prologue();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const PointI p = in[i];
    const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
    out[i] = make_float4((float)(M[0] * x + M[1] * y + M[2] * z + M[3]), (float)(M[4] * x + M[5] * y + M[6] * z + M[7]),
                         (float)(M[8] * x + M[9] * y + M[10] * z + M[11]), p.intensity);
  }
}

}  // namespace

extern "C" {

int floam_write_pcd_binary(const char* path, const floam_point_xyzi* pts, int n) {
  if (!path || n < 0 || (n > 0 && !pts)) return FLOAM_ERR_ARG;
  return write_pcd_xyzi(path, pts, (size_t)n);
}

// SavePosegraph, src/utils.cpp:3-79
int floam_save_posegraph(const char* directory, const double* poses16, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets, int n) {
  if (!args_ok(directory, poses16, stamps, clouds, offsets, n)) return FLOAM_ERR_ARG;
  const std::string dump_directory(directory);
  std::cout << "Save posegraph to:\n" << dump_directory << std::endl << std::endl;
  if (!make_dirs(dump_directory)) return FLOAM_ERR_ARG;
  std::ofstream graph_ofs(dump_directory + "/graph.g2o");
  for (int i = 0; i < n; ++i) {
    const double* M = poses16 + 16 * i;
    double R[9], q[4];
    rotation_of(M, R);
    m::quat_from_matrix(R, q);
    graph_ofs << "VERTEX_SE3:QUAT " << i << " " << M[3] << " " << M[7] << " " << M[11] << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << "\n";
  }
  graph_ofs << "FIX 0" << "\n";
  if (n <= 1) std::cerr << "cannot save a pose graph with only 1 vertex" << std::endl;
  for (int i = 0; i + 1 < n; ++i) {
    double rel[16], R[9], q[4];
    affine_between(poses16 + 16 * i, poses16 + 16 * (i + 1), rel);
    rotation_of(rel, R);
    m::quat_from_matrix(R, q);
    graph_ofs << "EDGE_SE3:QUAT " << i << " " << i + 1;
    graph_ofs << " " << rel[3] << " " << rel[7] << " " << rel[11] << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3];
    const double variances[6] = {0.01, 0.01, 0.01, 0.001, 0.001, 0.001};   // upper triangle of variances.asDiagonal()
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) graph_ofs << " " << (a == b ? variances[a] : 0.0);
    graph_ofs << "\n";
  }
  graph_ofs.close();
  for (int i = 0; i < n; ++i) {
    char sub[32];
    std::snprintf(sub, sizeof(sub), "/%06d", i);
    const std::string keyframe_directory = dump_directory + sub;
    if (!make_dirs(keyframe_directory)) return FLOAM_ERR_ARG;
    int rc = write_pcd_xyzi(keyframe_directory + "/cloud.pcd", clouds + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
    if (rc) return rc;
    unsigned int sec, nsec;
    ros_time(stamps[i], &sec, &nsec);
    std::ofstream data_ofs(keyframe_directory + "/data");
    data_ofs << "stamp " << sec << " " << nsec << "\n";
    data_ofs << "estimate\n"; print_matrix4(data_ofs, poses16 + 16 * i); data_ofs << "\n";
    data_ofs << "odom\n"; print_matrix4(data_ofs, poses16 + 16 * i); data_ofs << "\n";
    data_ofs << "accum_distance -1" << "\n";
    data_ofs << "id " << i << "\n";
  }
  return FLOAM_OK;
}

// SaveOdom, src/utils.cpp:82-106
int floam_save_odom(const char* directory, const double* poses16, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets, int n) {
  if (!args_ok(directory, poses16, stamps, clouds, offsets, n)) return FLOAM_ERR_ARG;
  const std::string dump_directory(directory);
  if (!make_dirs(dump_directory)) return FLOAM_ERR_ARG;
  std::cout << "Save odom to:\n" << dump_directory << std::endl << std::endl;
  for (int i = 0; i < n; ++i) {
    unsigned int sec, nsec;
    ros_time(stamps[i], &sec, &nsec);
    std::ostringstream name;   // boost::format("/%lf_%lf") % t.sec % t.nsec : integers streamed with the fixed flag set -> plain integers
    name << dump_directory << "/" << sec << "_" << nsec;
    const std::string filename = name.str();
    int rc = write_pcd_xyzi(filename + ".pcd", clouds + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
    if (rc) return rc;
    std::ofstream data_ofs(filename + ".odom");
    const double* M = poses16 + 16 * i;
    for (int r = 0; r < 4; ++r) data_ofs << M[r * 4] << " " << M[r * 4 + 1] << " " << M[r * 4 + 2] << " " << M[r * 4 + 3] << std::endl;
    data_ofs.close();
  }
  return FLOAM_OK;
}

// SavePosesHomogeneousBALM, src/odomEstimationNode.cpp:93-121
int floam_save_balm(const char* directory, const double* poses16, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets, int n) {
  if (!args_ok(directory, poses16, stamps, clouds, offsets, n)) return FLOAM_ERR_ARG;
  const std::string dir(directory);
  if (!make_dirs(dir)) return FLOAM_ERR_ARG;
  std::fstream stream((dir + "alidarPose.csv").c_str(), std::fstream::out);
  std::cout << "Save BALM to:\n" << dir << std::endl << std::endl;
  for (int i = 0; i < n; ++i) {
    // the node stores every cloud with header.stamp = toPCL(stamp) (microseconds) and reads it back through ros::Time (:102-104)
    unsigned int sec, nsec;
    ros_time(stamps[i], &sec, &nsec);
    const unsigned long long stamp_us = ((unsigned long long)sec * 1000000000ull + nsec) / 1000ull, ns = stamp_us * 1000ull;
    const double time = (double)(ns / 1000000000ull) + 1e-9 * (double)(ns % 1000000000ull);
    const double* m4 = poses16 + 16 * i;
    stream << std::fixed << m4[0] << "," << m4[1] << "," << m4[2] << "," << m4[3] << "," << std::endl
           << m4[4] << "," << m4[5] << "," << m4[6] << "," << m4[7] << "," << std::endl
           << m4[8] << "," << m4[9] << "," << m4[10] << "," << m4[11] << "," << std::endl
           << m4[12] << "," << m4[13] << "," << m4[14] << "," << time << "," << std::endl;
    int rc = write_pcd_xyzi(dir + "full" + std::to_string(i) + ".pcd", clouds + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
    if (rc) return rc;
  }
  return FLOAM_OK;
}

// SaveMerged, src/odomEstimationNode.cpp:66-92: every scan transformed by its pose and appended, the merged cloud saved, then
// downsampled with one VoxelGrid (leaf = downsample_size) and saved again when the result is not empty
int floam_save_merged(floam_ctx* c, const char* directory, const double* poses16, const floam_point_xyzi* clouds, const int64_t* offsets, int n,
                      double downsample_size) {
  if (!c || !directory || n < 0 || (n > 0 && (!poses16 || !offsets))) return FLOAM_ERR_ARG;
  if (c->inflight != 0) return FLOAM_ERR_ARG;
  const std::string dir(directory);
  if (!make_dirs(dir)) return FLOAM_ERR_ARG;
  std::cout << "Save merged point cloud to:\n" << dir << std::endl << std::endl;
  const long long total = n > 0 ? (long long)(offsets[n] - offsets[0]) : 0;
  if (total > c->stage_cap - 16) return FLOAM_ERR_CAPACITY;
  if (cudaSetDevice(c->device) != cudaSuccess) return FLOAM_ERR_CUDA;
  double* d_M = reinterpret_cast<double*>(c->d_stage_out + (size_t)c->stage_cap - 16);   // the pose of the scan being transformed: tail of the output staging buffer
  long long done = 0;
  for (int i = 0; i < n; ++i) {
    const int cnt = (int)(offsets[i + 1] - offsets[i]);
    if (cnt <= 0) continue;
    if (!clouds) return FLOAM_ERR_ARG;
    FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_stage_in, clouds + offsets[i], (size_t)cnt * 32, cudaMemcpyHostToDevice, c->stream));
    FLOAM_CUDA_OK(cudaMemcpyAsync(d_M, poses16 + 16 * i, 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int g = (cnt + 255) / 256;
    if (g > kNumSMs * 8) g = kNumSMs * 8;
    g_launches++;
    transform_affine_kernel<<<g, 256, 0, c->stream>>>((const PointI*)c->d_stage_in, cnt, d_M, c->d_stage_p4 + done);
    FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));   // the pageable pose / cloud sources are reused by the next iteration
    done += cnt;
  }
  std::cout << "Downsample point cloud resolution " << downsample_size << std::endl;
  std::vector<float> host((size_t)std::max<long long>(total, 1) * 4);
  if (total > 0) FLOAM_CUDA_OK(cudaMemcpy(host.data(), c->d_stage_p4, (size_t)total * 16, cudaMemcpyDeviceToHost));
  int rc = write_pcd(dir + "floam_merged.pcd", host.data(), (size_t)total);
  if (rc) return rc;
  int n_ds = 0;
  if (total > 0) {
    c->h_ints[0] = (int)total;
    FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_stage_n, &c->h_ints[0], 4, cudaMemcpyHostToDevice, c->stream));
    FLOAM_CUDA_OK(cudaMemsetAsync(c->d_stage_n + 1, 0, 4, c->stream));
    voxel_grid_device(c->d_stage_p4, 16, c->d_stage_n, (int)total, (float)downsample_size, c->d_stage_out, c->d_stage_n + 1, c->vws, nullptr, c->stream);
    FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 20, c->d_stage_n + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
    n_ds = c->h_ints[20];
    if (cudaGetLastError() != cudaSuccess) return FLOAM_ERR_CUDA;
  }
  if (n_ds > 0) {
    FLOAM_CUDA_OK(cudaMemcpy(host.data(), c->d_stage_out, (size_t)n_ds * 16, cudaMemcpyDeviceToHost));
    return write_pcd(dir + "floam_merged_downsampled_leaf_" + std::to_string(downsample_size) + ".pcd", host.data(), (size_t)n_ds);
  }
  std::cout << "No downsampled point cloud saved - increase \"output_downsample_size\"" << std::endl;
  return FLOAM_OK;
}

}  // extern "C"
