/* floam_b200 — C ABI of the B200-native FLOAM odometry hot path.
 *
 * The reference (dan11003/floam) has no FFI layer: its boundary is the C++ class API consumed by the three ROS nodes
 * (SURVEY.md §8b).  This header is the plain-C surface those classes are re-implemented on; the header-only C++ shims in
 * floam_b200/host/ give back the reference signatures (LaserProcessingClass, OdomEstimationClass, LaserMappingClass,
 * dmapping::ImuHandler / Compensate) on top of it.  Every entry point cites the reference interface it replaces.
 *
 * Conventions: all pointers are HOST pointers unless the name says otherwise; the context owns every device buffer, one
 * CUDA stream and one device; a context is not thread-safe (one context per sequence; contexts on different GPUs are
 * independent — multi-GPU = replicas, no collective).  Every call returns a status (0 = OK) and never throws.
 * There is no CPU fallback: without a CUDA device floam_create fails with FLOAM_ERR_NO_DEVICE.
 */
#ifndef FLOAM_B200_H_
#define FLOAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct floam_ctx floam_ctx;

/* vel_point::PointXYZIRT, reference include/lidar.h:14-32 (32 bytes, 16-aligned) */
typedef struct floam_point_xyzirt {
  float x, y, z, _pad0;
  float intensity;
  uint16_t ring, _pad1;
  float time;
  float _pad2;
} floam_point_xyzirt;

/* pcl::PointXYZI (32 bytes, intensity at offset 16) */
typedef struct floam_point_xyzi {
  float x, y, z, _pad0;
  float intensity;
  float _pad1[3];
} floam_point_xyzi;

enum floam_status {
  FLOAM_OK = 0,
  FLOAM_ERR_NO_DEVICE = 1,   /* no CUDA device / wrong architecture: the product path refuses to run */
  FLOAM_ERR_CUDA = 2,
  FLOAM_ERR_CAPACITY = 3,    /* an input or intermediate exceeded the capacities given at create time */
  FLOAM_ERR_ARG = 4,
  FLOAM_NO_IMU = 5,          /* dmapping::Compensate returned false ("no imu data"), reference src/dataHandler.cpp:99-102 */
  FLOAM_ERR_NONFINITE = 6    /* non-finite input coordinates (undefined in the reference, Q9) */
};

/* Error semantics of the frame path (floam_process_submit / _wait / _scan / _staged): FLOAM_ERR_NONFINITE and FLOAM_ERR_CAPACITY are
 * reported for the frame that raised them and only for that frame (the pose of such a frame is still computed and returned; the
 * reference has no error path at all and keeps running on NaN input, src/laserProcessingClass.cpp:74-75).  The next frame starts with
 * clean flags.  floam_replay_staged reports the OR over the frames it replayed.  FLOAM_ERR_CUDA is fatal for the context. */

enum floam_loss {
  FLOAM_LOSS_TRIVIAL = 0,    /* what the reference does for "cauchy" (loss_function stays nullptr, src/odomEstimationClass.cpp:88-91) */
  FLOAM_LOSS_HUBER = 1,      /* ceres::HuberLoss(0.1), src/odomEstimationClass.cpp:86 */
  FLOAM_LOSS_CAUCHY_TRUE = 2 /* opt-in: ceres::CauchyLoss(0.2) actually applied (not reachable in the reference) */
};

/* Opt-in algorithmic fixes.  All off by default: the default path reproduces the reference, quirks included.
 *  SINGLE_PREDICTION : the second pass of the deskew mode does not predict again.  The reference's test `update_type == VANILLA ||
 *                      UpdateType::INITIAL_ITERATION` is always true (src/odomEstimationClass.cpp:62-68, Q2), so pass 2 starts one
 *                      inter-frame motion beyond the pass-1 result and overwrites last_odom with it; with the fix pass 2 starts AT the
 *                      pass-1 result and last_odom stays the previous frame's pose, as the comment at :66 intends.
 *  ROTATED_VELOCITY  : dmapping::CompensateVelocity adds the world-frame velocity to sensor-frame points without rotating it
 *                      (src/dataHandler.cpp:82-91, include/odomEstimationClass.h:78, Q14); with the fix p += R(odom)^T v * t.
 *  IMU_SLERP         : ImuHandler::Get interpolates the two samples around the stamp with Eigen's slerp at
 *                      tSlerp = (t - t_before) / (t_after - t_before), which the reference computes and then ignores
 *                      (Interpolate returns data1, src/dataHandler.cpp:48-50,61-62), instead of the zero-order hold. */
enum floam_fix { FLOAM_FIX_SINGLE_PREDICTION = 1, FLOAM_FIX_ROTATED_VELOCITY = 2, FLOAM_FIX_IMU_SLERP = 4 };

enum floam_update_type { FLOAM_VANILLA = 0, FLOAM_INITIAL_ITERATION = 1, FLOAM_REFINEMENT_AND_UPDATE = 2 }; /* include/odomEstimationClass.h:63 */

/* lidar::Lidar (include/lidar.h:53-86) + OdomEstimationClass::init arguments + capacities */
typedef struct floam_params {
  int num_lines;            /* /scan_line, default 64 */
  double scan_period;       /* /scan_period, default 0.1 */
  double vertical_angle;    /* /vertical_angle, default 2.0 (unused by the algorithms) */
  double max_distance;      /* /max_dis, default 60 */
  double min_distance;      /* /min_dis, default 2 */
  double map_resolution;    /* /map_resolution, default 0.4 */
  int loss;                 /* floam_loss; floam_loss_from_string maps the reference's string */
  int max_scan_points;      /* capacity of one scan (default 300000) */
  int max_map_points;       /* capacity of each local map, edge and surf (default 4,000,000) */
  int max_global_map_points;/* capacity of the LaserMappingClass map (default 8,000,000; 0 = mapping disabled) */
  int max_grid_cells;       /* capacity of each local map's 1 m search grid, in cells (default 8,388,608) */
  int fixes;                /* OR of floam_fix bits; default 0 = reference behaviour */
} floam_params;

void floam_params_default(floam_params* p);
int floam_loss_from_string(const char* loss_function); /* lower-cases like src/odomEstimationClass.cpp:23; "huber" -> HUBER, else TRIVIAL */
const char* floam_status_string(int status);
const char* floam_version(void);

/* LaserProcessingClass::init + OdomEstimationClass::init + LaserMappingClass::init
 * (src/laserProcessingClass.cpp:6, src/odomEstimationClass.cpp:7-26, src/laserMappingClass.cpp:7-32) */
int floam_create(const floam_params* params, int device, floam_ctx** out);
void floam_destroy(floam_ctx* ctx);
/* page-locked host buffers for scans handed to floam_process_submit (so the upload overlaps the previous frame's kernels) */
void* floam_alloc_pinned(size_t bytes);
void floam_free_pinned(void* p);
/* per-frame launch sequences are replayed as CUDA graphs by default; 0 launches the kernels one by one (debugging / profiling) */
int floam_set_graphs(floam_ctx* ctx, int enabled);
/* addPointsToMap's CropBox + VoxelGrid (src/odomEstimationClass.cpp:270-292) either re-sorts map + new points, or sorts only the new
 * and the out-of-place points and merges them into the map, which is in voxel order already (faster from ~250k map points on).
 * Identical maps either way. mode 0 = always re-sort, 1 = by map size (default), 2 = always merge; for A/B measurements and tests.
 * No frame may be in flight. Process default: environment FLOAM_MAP_MERGE (and FLOAM_MAP_MERGE_MIN = the size threshold). */
int floam_set_map_merge(floam_ctx* ctx, int mode);

/* Quaternions cross this ABI in Eigen coefficient order (x, y, z, w), like parameters[0..3] of the reference. */
/* dmapping::ImuHandler::AddMsg (src/dataHandler.cpp:24-40): drops samples <= 10 us after the previous one. */
int floam_imu_push(floam_ctx* ctx, double stamp, const double q_xyzw[4]);
/* dmapping::ImuHandler::Get (src/dataHandler.cpp:51-75): zero-order hold; *valid = 0 and a zero quaternion when not covered. */
int floam_imu_get(floam_ctx* ctx, double stamp, double q_xyzw[4], int* valid);
int floam_imu_size(floam_ctx* ctx, int* n); /* ImuHandler::size() */
int floam_imu_time_contained(floam_ctx* ctx, double stamp, int* contained); /* ImuHandler::TimeContained (src/dataHandler.cpp:76-81) */

/* CenterTime + dmapping::Compensate + IMU alignment, in place (src/laserProcessingNode.cpp:65-78,108-116; src/dataHandler.cpp:93-122).
 * stamp_us is the pcl header stamp (microseconds) and is re-centred like the reference. Returns FLOAM_NO_IMU when
 * Compensate would return false (points are then only time-centred). */
int floam_deskew_align(floam_ctx* ctx, floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extrinsics_xyzw[4]);
/* The same pass with the three steps selectable, so that dmapping::Compensate alone (src/dataHandler.cpp:93-122) can be served. */
enum floam_deskew_flags { FLOAM_DESKEW_CENTER_TIME = 1, FLOAM_DESKEW_COMPENSATE = 2, FLOAM_DESKEW_ALIGN = 4 };
int floam_deskew_align_ex(floam_ctx* ctx, floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extrinsics_xyzw[4], int flags);
/* dmapping::CompensateVelocity (src/dataHandler.cpp:82-91), in place: p += velocity * p.time (no rotation, Q14) */
int floam_compensate_velocity(floam_ctx* ctx, floam_point_xyzirt* pts, int n, const double velocity[3]);

/* LaserProcessingClass::featureExtraction (src/laserProcessingClass.cpp:72-231). Appending is the caller's job:
 * edge/surf receive *ne / *ns points (capacities in points). */
int floam_feature_extract(floam_ctx* ctx, const floam_point_xyzirt* pts, int n,
                          floam_point_xyzirt* edge, int edge_cap, int* ne,
                          floam_point_xyzirt* surf, int surf_cap, int* ns);

/* OdomEstimationClass::initMapWithPoints (src/odomEstimationClass.cpp:28-32) */
int floam_odom_init_map(floam_ctx* ctx, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns);
/* OdomEstimationClass::UpdatePointsToMapSelector (src/odomEstimationClass.cpp:34-50). In deskew mode edge/surf are
 * velocity-compensated in place like the reference. pose_out = parameters[7] = (qx,qy,qz,qw,tx,ty,tz). */
int floam_odom_update(floam_ctx* ctx, floam_point_xyzirt* edge, int ne, floam_point_xyzirt* surf, int ns, int deskew, double pose_out[7]);
/* OdomEstimationClass::updatePointsToMap(PointXYZI overload, src/odomEstimationClass.cpp:57-124) */
int floam_odom_update_xyzi(floam_ctx* ctx, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns, int update_type, double pose_out[7]);
/* public member `odom` (Isometry3d, row-major 4x4) and GetVelocity() (include/odomEstimationClass.h:78,82) */
int floam_odom_get(floam_ctx* ctx, double odom_rowmajor[16], double velocity[3]);
/* public members laserCloudCornerMap / laserCloudSurfMap; getMap (src/odomEstimationClass.cpp:296-300) concatenates surf+edge */
int floam_odom_map_sizes(floam_ctx* ctx, int* n_edge, int* n_surf);
int floam_odom_get_map(floam_ctx* ctx, floam_point_xyzi* edge, int edge_cap, floam_point_xyzi* surf, int surf_cap);
/* test/stage-parity hooks: overwrite pose state / maps (no reference equivalent; the members are public or file-static there) */
int floam_odom_set_state(floam_ctx* ctx, const double odom_rowmajor[16], const double last_odom_rowmajor[16], int optimization_count);
int floam_odom_get_state(floam_ctx* ctx, double odom_rowmajor[16], double last_odom_rowmajor[16], int* optimization_count);
int floam_odom_set_map(floam_ctx* ctx, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns);

/* Fused, device-resident frame: scan -> features -> (first frame: initMapWithPoints, else UpdatePointsToMapSelector),
 * i.e. laser_processing() + odom_estimation() of src/laserProcessingNode.cpp:80-160 and src/odomEstimationNode.cpp:167-290
 * without the TCPROS hops. Only the scan goes up and only the 7-double pose comes down. */
int floam_process_scan(floam_ctx* ctx, const floam_point_xyzirt* pts, int n, int deskew, double pose_out[7]);
/* Same, split so the upload of frame k+1 overlaps the kernels of frame k. pts must stay valid (ideally pinned) until
 * the matching floam_process_wait returns. At most three submissions may be in flight (results come back in order). */
int floam_process_submit(floam_ctx* ctx, const floam_point_xyzirt* pts, int n, int deskew);
int floam_process_wait(floam_ctx* ctx, double pose_out[7]);
/* The same with the IMU steps of laser_processing() folded in: CenterTime + dmapping::Compensate + IMU alignment run on the
 * uploaded scan before feature extraction (src/laserProcessingNode.cpp:100-116). *stamp_us is re-centred like the reference.
 * Returns FLOAM_NO_IMU — and processes nothing — when Compensate would return false (the node skips such scans, :108-112). */
int floam_process_submit_imu(floam_ctx* ctx, const floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extrinsics_xyzw[4], int deskew);
int floam_process_scan_imu(floam_ctx* ctx, const floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extrinsics_xyzw[4], int deskew,
                           double pose_out[7]);
/* sensor_msgs/PointCloud2 ingestion: what pcl::fromROSMsg(*msg, *pointcloud_in) does on the host in the reference
 * (src/laserProcessingNode.cpp:98).  The raw message bytes (msg.data, row_step * height of them) are uploaded as they are — 22 bytes
 * per point for the Velodyne driver's XYZIRT layout instead of 32 — and re-packed into floam_point_xyzirt on the device.
 * off_* = byte offset inside a point of the message field with that NAME and the registered DATATYPE (FLOAT32; UINT16 for ring:
 * include/lidar.h:25-31), or -1 when the message has no such field; like fromROSMsg, an unmatched field stays 0. */
typedef struct floam_pc2_layout {
  uint32_t width, height;   /* msg.width, msg.height: n = width * height points */
  uint32_t point_step;      /* msg.point_step */
  uint32_t row_step;        /* msg.row_step (>= width * point_step) */
  int32_t off_x, off_y, off_z, off_intensity, off_ring, off_time;
  int32_t is_bigendian;     /* msg.is_bigendian */
} floam_pc2_layout;
/* stage entry point: unpack only (out has width * height elements) */
int floam_unpack_pointcloud2(floam_ctx* ctx, const uint8_t* data, const floam_pc2_layout* layout, floam_point_xyzirt* out);
/* floam_process_submit / floam_process_submit_imu fed with the raw message. stamp_us and extrinsics_xyzw are either both NULL
 * (no IMU steps) or both given (CenterTime + Compensate + alignment as in floam_process_submit_imu; FLOAM_NO_IMU drops the scan).
 * data must stay valid until the matching floam_process_wait returns. */
int floam_process_submit_pc2(floam_ctx* ctx, const uint8_t* data, const floam_pc2_layout* layout, uint64_t* stamp_us,
                             const double extrinsics_xyzw[4], int deskew);
/* Device-resident replay for kernel-only timing: the scans already sit in HBM (uploaded once with floam_stage_scans). */
int floam_stage_scans(floam_ctx* ctx, const floam_point_xyzirt* pts, const int64_t* offsets, int n_frames);
int floam_process_staged(floam_ctx* ctx, int frame, int deskew, double pose_out[7]);

/* Whole-sequence replay of staged frames [first, first + count): frames are enqueued back to back with no host round trip in
 * between (every decision of a frame is taken on the device); the 7-double poses come back from the device-side trajectory log. */
int floam_replay_staged(floam_ctx* ctx, int first, int count, int deskew, double* poses_out, float* total_ms);

/* LaserMappingClass::updateCurrentPointsToMap / getMap (src/laserMappingClass.cpp:148-200) */
int floam_mapping_update(floam_ctx* ctx, const floam_point_xyzi* pts, int n, const double pose_rowmajor[16]);
int floam_mapping_get_map(floam_ctx* ctx, floam_point_xyzi* out, int cap, int* n);
/* Incremental getMap() (the node republishes the whole map every frame, src/laserMappingNode.cpp:85-92): the points of
 * every 50 m cell that changed since the previous call — ALL points of such a cell, in the order getMap() lists them inside the cell —
 * each with its cell coordinates (cells_xyz: 3 ints per point). A caller that keeps one cloud per cell replaces the changed cells and
 * concatenates cells in (x, y, z) order to obtain exactly getMap()'s cloud (floam_b200/host/laserMappingClass.h does). out == NULL: only
 * *n is returned and the change marks are kept. */
int floam_mapping_get_changed_cells(floam_ctx* ctx, floam_point_xyzi* out, int32_t* cells_xyz, int cap, int* n);

/* On-disk outputs the odometry node writes when it exits (src/odomEstimationNode.cpp:66-121,373-387; src/utils.cpp:3-106).  Scans come as
 * one concatenated array of pcl::PointXYZI with offsets[n + 1], poses as row-major 4x4 matrices (Eigen::Affine3d::matrix()), stamps in
 * seconds.  Files are byte-identical to the reference's: iostream text formats, PCL 1.8 binary PCD (FIELDS x y z intensity). */
int floam_write_pcd_binary(const char* path, const floam_point_xyzi* pts, int n);               /* pcl::io::savePCDFileBinary<PointXYZI> */
int floam_save_posegraph(const char* directory, const double* poses16, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets,
                         int n);                                                                 /* SavePosegraph: graph.g2o + %06d/{cloud.pcd,data} */
int floam_save_odom(const char* directory, const double* poses16, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets,
                    int n);                                                                      /* SaveOdom: <sec>_<nsec>.pcd / .odom */
int floam_save_balm(const char* directory, const double* poses16, const double* stamps, const floam_point_xyzi* clouds, const int64_t* offsets,
                    int n);                                                                      /* SavePosesHomogeneousBALM: alidarPose.csv + full<i>.pcd (directory ends with '/') */
/* SaveMerged: scans transformed by their poses (pcl::transformPointCloud with an Affine3d), merged, saved, voxel-downsampled, saved again;
 * the transform and the VoxelGrid run on the device. directory ends with '/'. */
int floam_save_merged(floam_ctx* ctx, const char* directory, const double* poses16, const floam_point_xyzi* clouds, const int64_t* offsets, int n,
                      double downsample_size);

/* pcl::VoxelGrid<PointXYZI>::filter and pcl::CropBox<PointXYZI>::filter as used at src/odomEstimationClass.cpp:137-142,278-292
 * (stage entry points; also what floam_b200/host/mini_pcl.h's filters call) */
int floam_voxel_grid(floam_ctx* ctx, const floam_point_xyzi* pts, int n, float leaf, floam_point_xyzi* out, int cap, int* n_out);
int floam_crop_box(floam_ctx* ctx, const floam_point_xyzi* pts, int n, const float min_xyz[3], const float max_xyz[3], floam_point_xyzi* out, int cap, int* n_out);
/* The filter step of addPointsToMap (:270-292) on its own: VoxelGrid(CropBox(map_pts followed by new_pts)), through the merge path
 * the keyframe update uses (map_pts is expected to be mostly in voxel order; any input gives the exact filter result).
 * min_xyz / max_xyz both NULL: no crop box. */
int floam_voxel_grid_update(floam_ctx* ctx, const floam_point_xyzi* map_pts, int n_map, const floam_point_xyzi* new_pts, int n_new, float leaf,
                            const float min_xyz[3], const float max_xyz[3], floam_point_xyzi* out, int cap, int* n_out);
/* pcl::KdTreeFLANN::setInputCloud + nearestKSearch(k=5) (src/odomEstimationClass.cpp:78-79,153,206): exact ids for every
 * query whose 5th neighbour is closer than 1 m (the only ones the reference uses); others report ids = -1. */
int floam_knn5(floam_ctx* ctx, const floam_point_xyzi* map, int m, const floam_point_xyzi* queries, int nq, int* ids, float* sqdist);

/* Per-stage taps of the last odometry update (parity tests), see DESIGN.md for the record layouts. */
enum floam_debug_what {
  FLOAM_DBG_DS_EDGE = 0, FLOAM_DBG_DS_SURF = 1,       /* floam_point_xyzi[]  downsampled clouds */
  FLOAM_DBG_EDGE_KNN = 2, FLOAM_DBG_SURF_KNN = 3,     /* int[5*n]  ids of the last outer iteration (-1 = not accepted by the 1 m gate) */
  FLOAM_DBG_EDGE_D2 = 4, FLOAM_DBG_SURF_D2 = 5,       /* float[5*n] */
  FLOAM_DBG_EDGE_OK = 6, FLOAM_DBG_SURF_OK = 7,       /* uint8[n] residual accepted */
  FLOAM_DBG_RESIDUALS = 8,                            /* double[10*n] kind,curr(3),a(3),b(3) for accepted correspondences, edge then surf */
  FLOAM_DBG_LM = 9,                                   /* double[47]: iterations, accepted, initial_cost, final_cost, termination, H0[36], g0[6] */
  FLOAM_DBG_SCALARS = 10,                             /* int[6]: outer_iterations, keyframe, n_ds_edge, n_ds_surf, n_correspondences, solve_skipped */
  FLOAM_DBG_FEATURE_SRC_EDGE = 11, FLOAM_DBG_FEATURE_SRC_SURF = 12, /* int[]: input indices of the last feature extraction's outputs */
  FLOAM_DBG_CLOCKS = 13,                              /* int64[8]: SM clock stamps inside the last solve's final step attempt — profiling aid */
  FLOAM_DBG_TIMELINE = 14                             /* int64[8]: ns sums over frames {pose-dependent half, gap to the next one, solve part, frames} + last stamps — profiling aid */
};
int floam_debug_fetch(floam_ctx* ctx, int what, void* out, size_t cap_bytes, size_t* n_bytes);

/* Launch accounting for bench.py ("gpu_launches"): kernels launched by this context since the last reset. */
int floam_launch_count(floam_ctx* ctx, int64_t* launches, int reset);
/* Per-kernel-class device timing for the roofline leg of bench.py: while enabled, every kernel is bracketed by a CUDA event pair
 * on its stream; frame graphs are re-captured with the pairs as event-record nodes, so the measured durations contain no host
 * launch gaps. Slots are named by floam_kernel_name. */
int floam_set_kernel_timing(floam_ctx* ctx, int enabled);
int floam_kernel_slots(void);
const char* floam_kernel_name(int slot);
int floam_kernel_timing(floam_ctx* ctx, int slot, double* total_ms, int64_t* launches);
/* CUDA events around the device work of the last floam_process_* call, milliseconds. */
int floam_last_frame_ms(floam_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* FLOAM_B200_H_ */
