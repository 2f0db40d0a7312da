/*
 * mld_synth.h -- synthetic KITTI-shaped input for the benchmark and the tests. Not a reference
 * component: the reference ships no data. Deterministic (counter-based hashes of seed / frame / index); the host
 * generators (libmld_synth.so, plain C++, no CUDA) and the device generators (libmld_cuda.so,
 * mld_synth_*_device in mld_c_api.h) run the same inline model and agree bit for bit.
 *
 * Scene: one spinning lidar (rings x azimuth steps, point order azimuth-major then ring) 1.73 m above a ground
 * plane with random axis-aligned boxes, range noise and NaN dropouts. Features are integer pixel coordinates like
 * the ones tracklets_depth hands to the estimator (tracklets_depth/src/tracklet_depth_module.cpp:75-76), drawn from
 * four classes: on the sensor-facing faces of the boxes (corner detectors fire on objects, not on bare asphalt),
 * uniform inside the lidar-covered band, above the band (no lidar return: status 2), and -- for the road
 * configuration -- in the lower third of the image where the returns are ground returns.
 */
#ifndef MLD_SYNTH_H
#define MLD_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mld_synth_config {
    int32_t rings;            /* 64 (HDL-64) or 128 */
    int32_t azimuth_steps;    /* 1875 -> 120000 points */
    float elev_top_deg;       /* +2.0 */
    float elev_bottom_deg;    /* -24.8 */
    float sensor_height;      /* 1.73 m above ground */
    float max_range;          /* 120 m */
    float range_noise_sigma;  /* 0.02 m */
    float dropout_prob;       /* 0.02 -> NaN points */
    int32_t n_boxes;          /* <= 64 obstacles per frame */
    int32_t image_width, image_height;
    float band_top_frac;      /* fraction of the image height where the lidar-covered band starts */
    /* feature classes (fractions of the features of a frame; the rest is uniform inside the band) */
    float above_band_frac;    /* above the band: no lidar coverage */
    float object_frac;        /* on the sensor-facing face of a box (needs the camera below) */
    float road_frac;          /* lower third of the image (ground returns): the road configuration */
    /* camera used to place the object features: pinhole f, cx, cy and the row-major 3x4 lidar -> camera transform */
    float cam_f, cam_cx, cam_cy;
    float cam_T[12];
    int32_t two_block_rings;  /* 1: HDL-64E ring layout (upper half 1/3 degree apart, lower half ~1/2 degree), 0: uniform */
} mld_synth_config;

/* ---- host side: libmld_synth.so (no GPU, no CUDA runtime) ---- */
/* shape 0: KITTI (64 x 1875 points, 1241 x 376, KITTI calibration), 1: 128-beam (128 x 2032 points, 2048 x 1024);
 * road != 0 selects the road / non-road feature mix of BASELINE.json configs[2] */
void mld_synth_default_config(mld_synth_config* c, int dense);
void mld_synth_config_for(mld_synth_config* c, int dense, int road);
int64_t mld_synth_points_per_frame(const mld_synth_config* c);
/* points as float4 (x,y,z,intensity), features as 2 x F doubles (= Eigen::Matrix2Xd memory); 0 or -1 (bad argument) */
int mld_synth_points_host(const mld_synth_config* c, uint64_t seed, int64_t frame, float* out_xyzi);
int mld_synth_features_host(const mld_synth_config* c, uint64_t seed, int64_t frame, int F, double* out_uv);
/* same cloud as 32-byte pcl::PointXYZI records (x,y,z,pad,intensity,pad,pad,pad): the drop-in caller's layout */
int mld_synth_points_host_xyzi32(const mld_synth_config* c, uint64_t seed, int64_t frame, float* out_32b);

#ifdef __cplusplus
}
#endif
#endif
