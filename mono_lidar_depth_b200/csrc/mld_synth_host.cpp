// mld_synth_host.cpp -- libmld_synth.so: host generators of the synthetic KITTI-shaped input (include/mld_synth.h).
// Plain C++, no CUDA: the CPU arm of bench.py and the CPU tests generate their inputs here without ever mapping the
// product library. Built with -ffp-contract=off so that it agrees bit for bit with the device generators.
#include <vector>

#include "mld_synth_model.h"

extern "C" {

void mld_synth_config_for(mld_synth_config* c, int dense, int road) {
    memset(c, 0, sizeof(*c));
    c->rings = dense ? 128 : 64;
    c->azimuth_steps = dense ? 2032 : 1875;
    c->elev_top_deg = 2.0f;
    c->elev_bottom_deg = -24.8f;
    c->sensor_height = 1.73f;
    c->max_range = 120.0f;
    c->range_noise_sigma = 0.02f;
    c->dropout_prob = 0.02f;
    c->n_boxes = 40;
    c->two_block_rings = 1;
    c->image_width = dense ? 2048 : 1241;
    c->image_height = dense ? 1024 : 376;
    c->band_top_frac = 0.4f;
    // feature mix: calibrated with the oracle so that the status histogram under monolidar_fusion/parameters.yaml matches the
    // reference's own logs -- monolidar_fusion/Logs/log_depths.txt: 586 of 2009 features (29 %) got a depth;
    // Logs/log_depth_calc_stats.txt: 22.5 % success, 72.9 % insufficient points, 4.7 % no local maximum -- i.e. ~29 % Success,
    // ~60 % RadiusSearchInsufficientPoints, < 10 % HistogramNoLocalMax
    c->above_band_frac = road ? 0.2f : 0.58f;
    c->object_frac = road ? 0.25f : 0.31f;
    c->road_frac = road ? 0.5f : 0.0f;
    // KITTI raw 2011_09_26 calib_velo_to_cam (public calibration values) and the matching pinhole cameras
    c->cam_f = dense ? 1400.0f : 718.856f;
    c->cam_cx = dense ? 1024.0f : 607.1928f;
    c->cam_cy = dense ? 420.0f : 185.2157f;
    const float T[12] = {7.533745e-03f, -9.999714e-01f, -6.166020e-04f, -4.069766e-03f, 1.480249e-02f, 7.280733e-04f,
                         -9.998902e-01f, -7.631618e-02f, 9.998621e-01f, 7.523790e-03f, 1.480755e-02f, -2.717806e-01f};
    for (int i = 0; i < 12; i++) c->cam_T[i] = T[i];
}
void mld_synth_default_config(mld_synth_config* c, int dense) { mld_synth_config_for(c, dense, 0); }

int64_t mld_synth_points_per_frame(const mld_synth_config* c) { return (int64_t)c->rings * c->azimuth_steps; }

int mld_synth_points_host(const mld_synth_config* c, uint64_t seed, int64_t frame, float* out_xyzi) {
    if (!synth_config_ok(c) || !out_xyzi) return -1;
    std::vector<float> tables(synth_table_floats(*c));
    synth_build_tables(*c, tables.data());
    std::vector<SynthBox> boxes((size_t)c->n_boxes + 1);
    for (int b = 0; b < c->n_boxes; b++) boxes[(size_t)b] = synth_box(*c, seed, frame, b, tables.data());
    const long long n = (long long)c->rings * c->azimuth_steps;
    for (long long i = 0; i < n; i++) synth_point(*c, seed, frame, i, tables.data(), boxes.data(), out_xyzi + i * 4);
    return 0;
}

int mld_synth_points_host_xyzi32(const mld_synth_config* c, uint64_t seed, int64_t frame, float* out_32b) {
    if (!synth_config_ok(c) || !out_32b) return -1;
    const long long n = (long long)c->rings * c->azimuth_steps;
    std::vector<float> tmp((size_t)n * 4);
    if (mld_synth_points_host(c, seed, frame, tmp.data())) return -1;
    for (long long i = 0; i < n; i++) {  // pcl::PointXYZI: float data[4] (x,y,z,1) then intensity + 12 bytes of padding
        float* q = out_32b + i * 8;
        q[0] = tmp[(size_t)i * 4]; q[1] = tmp[(size_t)i * 4 + 1]; q[2] = tmp[(size_t)i * 4 + 2]; q[3] = 1.0f;
        q[4] = tmp[(size_t)i * 4 + 3]; q[5] = q[6] = q[7] = 0.0f;
    }
    return 0;
}

int mld_synth_features_host(const mld_synth_config* c, uint64_t seed, int64_t frame, int F, double* out_uv) {
    if (!synth_config_ok(c) || !out_uv || F < 0) return -1;
    std::vector<float> tables(synth_table_floats(*c));
    synth_build_tables(*c, tables.data());
    for (int i = 0; i < F; i++) synth_feature(*c, seed, frame, i, tables.data(), out_uv + (size_t)i * 2);
    return 0;
}

}  // extern "C"
