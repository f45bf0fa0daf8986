// Synthetic LiDAR sequence generator: procedurally generated scene, analytic ray casting,
// known ground-truth trajectory, optional motion distortion and a 200 Hz IMU orientation stream.
// This is workload tooling shared by tests and bench.py; it is neither the product path nor the oracle.
// Everything is a pure function of (seed, frame, ring, azimuth), so any thread count yields identical bytes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <vector>

namespace {

struct PointXYZIRT {  // byte-identical to vel_point::PointXYZIRT (reference include/lidar.h:14-32)
  float x, y, z, pad0;
  float intensity;
  std::uint16_t ring, pad1;
  float time;
  float pad2;
};
static_assert(sizeof(PointXYZIRT) == 32, "layout");

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }

inline std::uint64_t splitmix(std::uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline std::uint64_t key(std::uint64_t seed, std::uint64_t a, std::uint64_t b, std::uint64_t c, std::uint64_t d) {
  return splitmix(splitmix(splitmix(splitmix(seed ^ 0xF10A3000ull) + a) + b * 0x100000001B3ull) + c * 0x9E3779B1ull + d);
}
inline double u01(std::uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Box { double x0, y0, x1, y1, h; };
struct Cyl { double cx, cy, r, h; };

struct Pose { double R[3][3]; V3 t; };

struct Scene {
  std::uint64_t seed;
  int sensor;  // 0 VLP-16, 1 HDL-64E, 2 OS1-128
  int n_rings, n_az;
  double sigma;
  int distort;
  double speed;        // m/s along the road
  double scan_period;  // 0.1
  std::vector<double> elev;
};

// Endless street: the road centre-line is y_c(x) = 10 sin(x/50); the vehicle drives along +x and never revisits a place,
// so odometry drift cannot turn into loop-closure inconsistency. Buildings (axis-aligned boxes with half-columns on their
// faces) line both sides in 12 m lots, poles stand along the kerbs every 6 m; all derived from hashes of the lot index.
inline double road_y(double x) { return 10.0 * std::sin(x / 50.0); }
inline double road_dy(double x) { return 0.2 * std::cos(x / 50.0); }

Pose pose_at(const Scene& s, double t) {
  // start from rest: speed ramps up with a smoothstep over Ta seconds, then cruises (arclength in closed form)
  const double Ta = 4.0;
  double arc;
  if (t < Ta) { const double u = t / Ta; arc = s.speed * Ta * (u * u * u - 0.5 * u * u * u * u); }
  else arc = s.speed * Ta * 0.5 + s.speed * (t - Ta);
  const double px = arc, py = road_y(arc);
  const double yaw = std::atan2(road_dy(arc), 1.0);
  const double roll = 0.010 * std::sin(0.7 * t), pitch = 0.008 * std::sin(0.9 * t + 1.0);
  const double z = 1.73 + 0.02 * std::sin(1.3 * t);
  const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll), sr = std::sin(roll);
  Pose P;  // R = Rz(yaw) Ry(pitch) Rx(roll)
  P.R[0][0] = cy * cp; P.R[0][1] = cy * sp * sr - sy * cr; P.R[0][2] = cy * sp * cr + sy * sr;
  P.R[1][0] = sy * cp; P.R[1][1] = sy * sp * sr + cy * cr; P.R[1][2] = sy * sp * cr - cy * sr;
  P.R[2][0] = -sp;     P.R[2][1] = cp * sr;                P.R[2][2] = cp * cr;
  P.t = {px, py, z};
  return P;
}

void quat_from_R(const double R[3][3], double q[4]) {  // x,y,z,w
  double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0) * 2;
    q[3] = 0.25 * s; q[0] = (R[2][1] - R[1][2]) / s; q[1] = (R[0][2] - R[2][0]) / s; q[2] = (R[1][0] - R[0][1]) / s;
  } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
    double s = std::sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2;
    q[3] = (R[2][1] - R[1][2]) / s; q[0] = 0.25 * s; q[1] = (R[0][1] + R[1][0]) / s; q[2] = (R[0][2] + R[2][0]) / s;
  } else if (R[1][1] > R[2][2]) {
    double s = std::sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2;
    q[3] = (R[0][2] - R[2][0]) / s; q[0] = (R[0][1] + R[1][0]) / s; q[1] = 0.25 * s; q[2] = (R[1][2] + R[2][1]) / s;
  } else {
    double s = std::sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2;
    q[3] = (R[1][0] - R[0][1]) / s; q[0] = (R[0][2] + R[2][0]) / s; q[1] = (R[1][2] + R[2][1]) / s; q[2] = 0.25 * s;
  }
}

void build_scene(Scene& s) {
  s.elev.resize(s.n_rings);
  for (int i = 0; i < s.n_rings; ++i) {
    double deg;
    if (s.sensor == 0) deg = -15.0 + 2.0 * i;                                           // VLP-16
    else if (s.sensor == 1) deg = (i < 32) ? (2.0 - i / 3.0) : (-8.83 - (i - 32) / 2.0);  // HDL-64E (inverse of reference RingExtraction :50-56)
    else deg = -22.5 + 45.0 * i / (s.n_rings - 1);                                      // OS1-128
    s.elev[i] = deg * M_PI / 180.0;
  }
}

const double kView = 125.0;   // primitives farther than this along the road are never within the 120 m sensor range
const int kBins = 720;        // world-azimuth bins around the sensor (0.5 deg)

// Per-frame view of the scene: primitives near the sensor, bucketed by the world azimuth under which the sensor sees them.
struct FrameScene {
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
  std::vector<std::vector<int>> box_bins, cyl_bins;
};

void add_range(std::vector<std::vector<int>>& bins, double a0, double a1, int id) {  // a0 <= a1 (radians, may exceed [-pi,pi])
  int b0 = (int)std::floor((a0 + M_PI) / (2 * M_PI) * kBins), b1 = (int)std::floor((a1 + M_PI) / (2 * M_PI) * kBins);
  if (b1 - b0 >= kBins - 1) { for (int b = 0; b < kBins; ++b) bins[b].push_back(id); return; }
  for (int b = b0; b <= b1; ++b) bins[((b % kBins) + kBins) % kBins].push_back(id);
}

void build_frame_scene(const Scene& s, V3 o, FrameScene& fs) {
  fs.boxes.clear(); fs.cyls.clear();
  fs.box_bins.assign(kBins, std::vector<int>()); fs.cyl_bins.assign(kBins, std::vector<int>());
  const long k0 = (long)std::floor((o.x - kView) / 12.0), k1 = (long)std::floor((o.x + kView) / 12.0);
  for (long k = k0; k <= k1; ++k)
    for (int side = 0; side < 2; ++side) {
      const std::uint64_t kk = (std::uint64_t)(k + 1000000), sd = side;
      if (u01(key(s.seed, 3, kk, sd, 0)) > 0.88) continue;  // empty lot / side street
      const double wx = 5 + 6 * u01(key(s.seed, 3, kk, sd, 1)), wy = 5 + 7 * u01(key(s.seed, 3, kk, sd, 2));
      const double h = 3 + 12 * u01(key(s.seed, 3, kk, sd, 3));
      const double setback = 6 + 4 * u01(key(s.seed, 3, kk, sd, 4));
      const double cx = 12.0 * k + 6.0 + (u01(key(s.seed, 3, kk, sd, 5)) - 0.5) * (11.0 - wx);
      const double sgn = side ? 1.0 : -1.0;
      const double cy = road_y(cx) + sgn * (setback + wy / 2);
      Box b{cx - wx / 2, cy - wy / 2, cx + wx / 2, cy + wy / 2, h};
      fs.boxes.push_back(b);
      // facade detail: half-columns against every face (real vertical edges)
      const std::uint64_t bi = kk * 2 + sd;
      for (int face = 0; face < 4; ++face) {
        double len = (face < 2) ? wx : wy;
        double pos = 0.6 + 1.2 * u01(key(s.seed, 5, bi, face, 0));
        int c = 0;
        while (pos < len - 0.6) {
          double r = 0.12 + 0.13 * u01(key(s.seed, 5, bi, face, 1 + 2 * c));
          double px, py;
          if (face == 0) { px = b.x0 + pos; py = b.y0; }
          else if (face == 1) { px = b.x0 + pos; py = b.y1; }
          else if (face == 2) { px = b.x0; py = b.y0 + pos; }
          else { px = b.x1; py = b.y0 + pos; }
          fs.cyls.push_back({px, py, r, h});
          pos += 1.8 + 1.8 * u01(key(s.seed, 5, bi, face, 2 + 2 * c));
          ++c;
        }
      }
    }
  const long j0 = (long)std::floor((o.x - kView) / 6.0), j1 = (long)std::floor((o.x + kView) / 6.0);
  for (long j = j0; j <= j1; ++j)
    for (int side = 0; side < 2; ++side) {
      const std::uint64_t jj = (std::uint64_t)(j + 1000000), sd = side;
      if (u01(key(s.seed, 4, jj, sd, 0)) > 0.7) continue;
      const double px = 6.0 * j + 6.0 * u01(key(s.seed, 4, jj, sd, 1));
      const double py = road_y(px) + (side ? 1.0 : -1.0) * (3.5 + 1.5 * u01(key(s.seed, 4, jj, sd, 2)));
      fs.cyls.push_back({px, py, 0.1 + 0.2 * u01(key(s.seed, 4, jj, sd, 3)), 4 + 6 * u01(key(s.seed, 4, jj, sd, 4))});
    }
  // bucket by azimuth, padded for the sensor's own motion during one distorted scan (<= 1.5 m)
  for (int i = 0; i < (int)fs.boxes.size(); ++i) {
    const Box& b = fs.boxes[i];
    const double xs[2] = {b.x0, b.x1}, ys[2] = {b.y0, b.y1};
    if (o.x > b.x0 - 2 && o.x < b.x1 + 2 && o.y > b.y0 - 2 && o.y < b.y1 + 2) { add_range(fs.box_bins, -M_PI, M_PI, i); continue; }
    const double ac = std::atan2((b.y0 + b.y1) / 2 - o.y, (b.x0 + b.x1) / 2 - o.x);
    double lo = 0, hi = 0, dmin = 1e30;
    for (int a = 0; a < 2; ++a)
      for (int c = 0; c < 2; ++c) {
        double d = std::atan2(ys[c] - o.y, xs[a] - o.x) - ac;
        while (d > M_PI) d -= 2 * M_PI;
        while (d < -M_PI) d += 2 * M_PI;
        lo = std::min(lo, d); hi = std::max(hi, d);
        dmin = std::min(dmin, std::hypot(xs[a] - o.x, ys[c] - o.y));
      }
    const double cdx = std::max(std::max(b.x0 - o.x, o.x - b.x1), 0.0), cdy = std::max(std::max(b.y0 - o.y, o.y - b.y1), 0.0);
    dmin = std::max(std::min(dmin, std::hypot(cdx, cdy)), 0.5);
    const double pad = std::min(1.5 / dmin, 1.0) + 0.01;
    add_range(fs.box_bins, ac + lo - pad, ac + hi + pad, i);
  }
  for (int i = 0; i < (int)fs.cyls.size(); ++i) {
    const Cyl& c = fs.cyls[i];
    const double d = std::max(std::hypot(c.cx - o.x, c.cy - o.y), 0.5);
    const double ac = std::atan2(c.cy - o.y, c.cx - o.x);
    const double half = std::asin(std::min(c.r / d, 1.0)) + std::min(1.5 / d, 1.0) + 0.01;
    add_range(fs.cyl_bins, ac - half, ac + half, i);
  }
}

// nearest hit distance along o + t d (|d| = 1), or <0 when nothing is hit
double cast(const FrameScene& fs, V3 o, V3 d) {
  double best = 1e30;
  if (d.z < -1e-9) { double t = -o.z / d.z; if (t > 0 && t < best) best = t; }
  int bin = (int)std::floor((std::atan2(d.y, d.x) + M_PI) / (2 * M_PI) * kBins);
  bin = std::min(std::max(bin, 0), kBins - 1);
  const double ix = 1.0 / d.x, iy = 1.0 / d.y, iz = 1.0 / d.z;
  for (int id : fs.box_bins[bin]) {
    const Box& b = fs.boxes[id];
    double t0x = (b.x0 - o.x) * ix, t1x = (b.x1 - o.x) * ix; if (t0x > t1x) std::swap(t0x, t1x);
    double t0y = (b.y0 - o.y) * iy, t1y = (b.y1 - o.y) * iy; if (t0y > t1y) std::swap(t0y, t1y);
    double t0z = (0.0 - o.z) * iz, t1z = (b.h - o.z) * iz;   if (t0z > t1z) std::swap(t0z, t1z);
    double tn = std::max(t0x, std::max(t0y, t0z)), tf = std::min(t1x, std::min(t1y, t1z));
    if (tn <= tf && tn > 0 && tn < best) best = tn;
  }
  const double a = d.x * d.x + d.y * d.y;
  if (a > 1e-12) {
    for (int id : fs.cyl_bins[bin]) {
      const Cyl& c = fs.cyls[id];
      double ox = o.x - c.cx, oy = o.y - c.cy;
      double bq = ox * d.x + oy * d.y, cq = ox * ox + oy * oy - c.r * c.r;
      double disc = bq * bq - a * cq;
      if (disc < 0) continue;
      double t = (-bq - std::sqrt(disc)) / a;
      if (t > 0 && t < best) { double z = o.z + t * d.z; if (z >= 0 && z <= c.h) best = t; }
    }
  }
  return best < 1e29 ? best : -1.0;
}

int gen_scan(const Scene& s, int frame, PointXYZIRT* out, int cap) {
  const double t_frame = frame * s.scan_period;
  Pose P0 = pose_at(s, t_frame);
  FrameScene fs;
  build_frame_scene(s, P0.t, fs);
  int n = 0;
  for (int az = 0; az < s.n_az; ++az) {
    const double frac = (double)az / s.n_az;
    Pose P = s.distort ? pose_at(s, t_frame + frac * s.scan_period) : P0;
    const double phi = 2 * M_PI * frac;
    const double cphi = std::cos(phi), sphi = std::sin(phi);
    for (int ring = 0; ring < s.n_rings; ++ring) {
      const double ce = std::cos(s.elev[ring]), se = std::sin(s.elev[ring]);
      V3 ds{ce * cphi, ce * sphi, se};
      V3 dw{P.R[0][0] * ds.x + P.R[0][1] * ds.y + P.R[0][2] * ds.z, P.R[1][0] * ds.x + P.R[1][1] * ds.y + P.R[1][2] * ds.z,
            P.R[2][0] * ds.x + P.R[2][1] * ds.y + P.R[2][2] * ds.z};
      double r = cast(fs, P.t, dw);
      if (r < 0 || r > 120.0) continue;
      if (s.sigma > 0) {
        double u1 = u01(key(s.seed, 10, frame, ring, az)), u2 = u01(key(s.seed, 11, frame, ring, az));
        r += s.sigma * std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * M_PI * u2);
      }
      if (r < 0.3) continue;
      if (n >= cap) return n;
      PointXYZIRT p;
      std::memset(&p, 0, sizeof(p));
      p.x = (float)(r * ds.x); p.y = (float)(r * ds.y); p.z = (float)(r * ds.z); p.pad0 = 1.0f;
      p.intensity = (float)u01(key(s.seed, 12, frame, ring, az));
      p.ring = (std::uint16_t)ring;
      p.time = (float)(frac * s.scan_period);
      out[n++] = p;
    }
  }
  return n;
}

}  // namespace

extern "C" {

void* synth_create(unsigned long long seed, int sensor, int n_az, double sigma, int distort, double speed) {
  Scene* s = new Scene();
  s->seed = seed; s->sensor = sensor;
  s->n_rings = sensor == 0 ? 16 : (sensor == 1 ? 64 : 128);
  s->n_az = n_az; s->sigma = sigma; s->distort = distort; s->speed = speed; s->scan_period = 0.1;
  build_scene(*s);
  return s;
}
void synth_destroy(void* h) { delete (Scene*)h; }
int synth_num_rings(void* h) { return ((Scene*)h)->n_rings; }
int synth_max_points(void* h) { return ((Scene*)h)->n_rings * ((Scene*)h)->n_az; }
// ground-truth sensor pose (row-major 4x4) at time t seconds
void synth_pose(void* h, double t, double T[16]) {
  Pose P = pose_at(*(Scene*)h, t);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = P.R[i][j]; }
  T[3] = P.t.x; T[7] = P.t.y; T[11] = P.t.z; T[12] = T[13] = T[14] = 0; T[15] = 1;
}
// IMU orientation sample q_imu(t) = q_world_sensor(t) * extr^-1 with extr = yaw 180 deg (reference src/laserProcessingNode.cpp:196)
void synth_imu(void* h, double t, double q_xyzw[4]) {
  Pose P = pose_at(*(Scene*)h, t);
  double q[4];
  quat_from_R(P.R, q);
  // extr = (0,0,1,0) [x,y,z,w] ; extr^-1 = (0,0,-1,0) ; q * extr^-1
  const double ex = 0, ey = 0, ez = -1, ew = 0;
  q_xyzw[0] = q[3] * ex + q[0] * ew + q[1] * ez - q[2] * ey;
  q_xyzw[1] = q[3] * ey + q[1] * ew + q[2] * ex - q[0] * ez;
  q_xyzw[2] = q[3] * ez + q[2] * ew + q[0] * ey - q[1] * ex;
  q_xyzw[3] = q[3] * ew - q[0] * ex - q[1] * ey - q[2] * ez;
}
int synth_scan(void* h, int frame, void* out, int cap) { return gen_scan(*(Scene*)h, frame, (PointXYZIRT*)out, cap); }
// frames [frame0, frame0+n) into out (capacity cap_per_frame each, densely packed afterwards); counts[i] = points of frame i.
// Returns total points. Frames are generated on `threads` host threads.
long long synth_scans(void* h, int frame0, int n, void* out, int cap_per_frame, int* counts, int threads) {
  Scene& s = *(Scene*)h;
  PointXYZIRT* base = (PointXYZIRT*)out;
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; ++w)
    pool.emplace_back([&, w]() {
      for (int i = w; i < n; i += threads) counts[i] = gen_scan(s, frame0 + i, base + (size_t)i * cap_per_frame, cap_per_frame);
    });
  for (auto& t : pool) t.join();
  long long total = 0;
  for (int i = 0; i < n; ++i) {  // compact in place
    if (total != (long long)i * cap_per_frame) std::memmove(base + total, base + (size_t)i * cap_per_frame, sizeof(PointXYZIRT) * (size_t)counts[i]);
    total += counts[i];
  }
  return total;
}

}  // extern "C"
