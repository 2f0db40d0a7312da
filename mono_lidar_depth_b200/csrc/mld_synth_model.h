// mld_synth_model.h -- the synthetic scene of include/mld_synth.h as inline functions that compile both as plain C++
// (libmld_synth.so, host generators) and as CUDA (mld_synth.cu, device generators).
//
// Host and device run the SAME code: only exactly rounded float operations (+ - * /, comparisons via ternaries) on
// hashes and on trig tables computed once on the host, no multiply-add contraction on either side
// (-fmad=false / -ffp-contract=off), so both produce identical bits.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "mld_hash.h"
#include "mld_synth.h"

struct SynthBox {
    float lox, hix, loy, hiy, loz, hiz;
};
constexpr int SYNTH_MAX_BOXES = 64;

MLD_HD float synth_u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }
MLD_HD float synth_u01b(uint64_t h) { return (float)((h >> 16) & 0xffffffull) * (1.0f / 16777216.0f); }
MLD_HD float synth_qnan() {
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7fc00000);
#else
    uint32_t b = 0x7fc00000u;
    float f;
    memcpy(&f, &b, sizeof(f));
    return f;
#endif
}
MLD_HD float synth_min(float a, float b) { return (a < b) ? a : b; }
MLD_HD float synth_max(float a, float b) { return (a > b) ? a : b; }

// tables: cos/sin of the ring elevations (2 * rings), cos/sin of the azimuth steps (2 * azimuth_steps)
inline size_t synth_table_floats(const mld_synth_config& c) { return (size_t)(2 * c.rings + 2 * c.azimuth_steps); }
inline void synth_build_tables(const mld_synth_config& c, float* tables) {
    const double deg = 3.14159265358979323846 / 180.0;
    for (int e = 0; e < c.rings; e++) {
        double el;
        const int half = c.rings / 2;
        if (c.two_block_rings && half >= 2) {
            // HDL-64E layout: the upper half of the lasers is spaced 1/3 degree, the lower half ~1/2 degree (64 rings: +2 ... -8.33
            // and -8.83 ... elev_bottom); a 128-ring sensor halves both spacings
            const double upper_span = 31.0 / 3.0, gap = 0.5 * 64.0 / (double)c.rings;
            const double lower_top = (double)c.elev_top_deg - upper_span - gap;
            if (e < half)
                el = (double)c.elev_top_deg - upper_span * (double)e / (double)(half - 1);
            else
                el = lower_top + ((double)c.elev_bottom_deg - lower_top) * (double)(e - half) / (double)(c.rings - half - 1);
        } else {
            el = (double)c.elev_top_deg +
                 ((double)c.elev_bottom_deg - (double)c.elev_top_deg) * (c.rings > 1 ? (double)e / (double)(c.rings - 1) : 0.0);
        }
        tables[e] = (float)cos(el * deg);
        tables[c.rings + e] = (float)sin(el * deg);
    }
    for (int a = 0; a < c.azimuth_steps; a++) {
        // azimuth 0 looks along +x; the sweep starts behind the sensor so that frontal points are mid-cloud
        double az = -180.0 + 360.0 * (double)a / (double)c.azimuth_steps;
        tables[2 * c.rings + a] = (float)cos(az * deg);
        tables[2 * c.rings + c.azimuth_steps + a] = (float)sin(az * deg);
    }
}
inline bool synth_config_ok(const mld_synth_config* c) {
    return c && c->rings > 0 && c->azimuth_steps >= 6 && c->n_boxes >= 0 && c->n_boxes <= SYNTH_MAX_BOXES && c->image_width > 0 &&
           c->image_height > 1;
}

MLD_HD SynthBox synth_box(const mld_synth_config& c, uint64_t seed, long long frame, int b, const float* tables) {
    const float* cos_az = tables + 2 * c.rings;
    const float* sin_az = cos_az + c.azimuth_steps;
    uint64_t key = seed ^ 0xB0C5B0C5ull;
    uint64_t h0 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 0);
    uint64_t h1 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 1);
    uint64_t h2 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 2);
    uint64_t h3 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 3);
    uint64_t h4 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 4);
    uint64_t h5 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 5);
    // a quarter of the boxes stands inside the camera's field of view (+-35 degrees) at 12-45 m, where the upper block of the
    // lasers samples a face densely (every street scene has structure in view); of the rest two thirds sit in the +-60 degree
    // sector in front of the sensor (the camera looks along +x), one third anywhere around it
    int az;
    float r;
    if (4 * b < c.n_boxes) {
        const int sector = (c.azimuth_steps * 35) / 360;
        const int off = (int)(h1 % (uint64_t)(2 * sector + 1)) - sector;
        az = (c.azimuth_steps / 2 + off + c.azimuth_steps) % c.azimuth_steps;  // table index steps / 2 is azimuth 0 = straight ahead
        r = 12.0f + 33.0f * synth_u01(h2);
    } else {
        const int sector = c.azimuth_steps / 6;
        if ((h0 % 3ull) != 0ull) {
            const int off = (int)(h1 % (uint64_t)(2 * sector + 1)) - sector;
            az = (c.azimuth_steps / 2 + off + c.azimuth_steps) % c.azimuth_steps;
        } else {
            az = (int)(h1 % (uint64_t)c.azimuth_steps);
        }
        r = 8.0f + 72.0f * synth_u01(h2) * synth_u01b(h2);  // denser near the sensor, never on top of it
    }
    float cx = r * cos_az[az], cy = r * sin_az[az];
    float hx = 0.5f + 2.5f * synth_u01(h3), hy = 0.5f + 2.5f * synth_u01(h4), hh = 0.5f + 3.5f * synth_u01(h5);
    SynthBox bx;
    bx.lox = cx - hx; bx.hix = cx + hx;
    bx.loy = cy - hy; bx.hiy = cy + hy;
    bx.loz = -c.sensor_height; bx.hiz = -c.sensor_height + hh;
    return bx;
}

// range to the nearest surface along direction (dx,dy,dz) or a negative value for "no return"
MLD_HD float synth_cast(const mld_synth_config& c, float dx, float dy, float dz, const SynthBox* boxes, int nb) {
    float best = c.max_range;
    bool hit = false;
    if (dz < 0.0f) {
        float t = (-c.sensor_height) / dz;
        if (t < best) {
            best = t;
            hit = true;
        }
    }
    for (int b = 0; b < nb; b++) {
        const SynthBox& bx = boxes[b];
        float tx1 = bx.lox / dx, tx2 = bx.hix / dx;
        float ty1 = bx.loy / dy, ty2 = bx.hiy / dy;
        float tz1 = bx.loz / dz, tz2 = bx.hiz / dz;
        float tn = synth_max(synth_max(synth_min(tx1, tx2), synth_min(ty1, ty2)), synth_min(tz1, tz2));
        float tf = synth_min(synth_min(synth_max(tx1, tx2), synth_max(ty1, ty2)), synth_max(tz1, tz2));
        if (tn <= tf && tn > 0.5f && tn < best) {
            best = tn;
            hit = true;
        }
    }
    return hit ? best : -1.0f;
}

MLD_HD void synth_point(const mld_synth_config& c, uint64_t seed, long long frame, long long idx, const float* tables,
                        const SynthBox* boxes, float out[4]) {
    const float* cos_el = tables;
    const float* sin_el = tables + c.rings;
    const float* cos_az = tables + 2 * c.rings;
    const float* sin_az = cos_az + c.azimuth_steps;
    int a = (int)(idx / c.rings), e = (int)(idx % c.rings);
    float dx = cos_el[e] * cos_az[a], dy = cos_el[e] * sin_az[a], dz = sin_el[e];
    uint64_t key = seed ^ 0x9017C10Dull;
    uint64_t hd = mld_hash3(key, (uint64_t)frame, (uint64_t)idx, 0);
    uint64_t hn = mld_hash3(key, (uint64_t)frame, (uint64_t)idx, 1);
    const float qnan = synth_qnan();
    float t = synth_cast(c, dx, dy, dz, boxes, c.n_boxes);
    bool drop = synth_u01(hd) < c.dropout_prob;
    if (t < 0.0f || drop) {
        out[0] = qnan; out[1] = qnan; out[2] = qnan; out[3] = 0.0f;
        return;
    }
    // Irwin-Hall(4) noise, unit variance after scaling by sqrt(3)
    float s = (float)(hn & 0xffff) * (1.0f / 65536.0f) + (float)((hn >> 16) & 0xffff) * (1.0f / 65536.0f) +
              (float)((hn >> 32) & 0xffff) * (1.0f / 65536.0f) + (float)((hn >> 48) & 0xffff) * (1.0f / 65536.0f);
    float noise = (s - 2.0f) * 1.7320508f * c.range_noise_sigma;
    float tr = t + noise;
    out[0] = dx * tr; out[1] = dy * tr; out[2] = dz * tr;
    out[3] = synth_u01b(hd);
}

// a feature on the sensor-facing (x = lox) face of a box in front of the camera, inside the lidar's vertical field
// of view; false when no box of the frame offers one
MLD_HD bool synth_object_feature(const mld_synth_config& c, uint64_t seed, long long frame, int i, const float* tables, int& u,
                                 int& v) {
    if (c.n_boxes <= 0 || !(c.cam_f > 0.0f)) return false;
    const uint64_t key = seed ^ 0x0B1EC7F5ull;
    const float tan_top = 0.0262f;  // returns exist up to ~1.5 degrees above the horizon (elev_top is +2)
    // the boxes are tried cyclically from a hashed start until one offers a usable face
    const uint64_t ha = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 0);
    const uint64_t hb = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 1);
    const int first = (int)(ha % (uint64_t)c.n_boxes);
    for (int attempt = 0; attempt < c.n_boxes; attempt++) {
        int b = first + attempt;
        if (b >= c.n_boxes) b -= c.n_boxes;
        const SynthBox bx = synth_box(c, seed, frame, b, tables);
        // faces between 12 and 60 m are sampled by the densely spaced upper block of the lasers; stay off the face's edges
        if (!(bx.lox > 12.0f) || !(bx.lox < 60.0f)) continue;
        const float ztop = synth_min(bx.hiz, bx.lox * tan_top) - 0.15f, zbot = bx.loz + 0.3f;
        if (!(ztop > zbot)) continue;
        const float x = bx.lox;
        const float y = bx.loy + (bx.hiy - bx.loy) * (0.1f + 0.8f * synth_u01(hb));
        const float z = zbot + (ztop - zbot) * synth_u01b(hb);
        const float* T = c.cam_T;
        const float X = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        const float Y = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        const float Z = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
        if (!(Z > 1.0f)) continue;
        const float fu = c.cam_f * X / Z + c.cam_cx, fv = c.cam_f * Y / Z + c.cam_cy;
        if (!(fu >= 0.0f) || !(fv >= 0.0f) || !(fu < (float)c.image_width) || !(fv < (float)c.image_height)) continue;
        u = (int)fu;
        v = (int)fv;
        return true;
    }
    return false;
}

MLD_HD void synth_feature(const mld_synth_config& c, uint64_t seed, long long frame, int i, const float* tables, double out[2]) {
    uint64_t key = seed ^ 0xFEA7FEA7ull;
    uint64_t h0 = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 0);
    uint64_t h1 = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 1);
    uint64_t h2 = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 2);
    int W = c.image_width, H = c.image_height;
    int band_top = (int)(c.band_top_frac * (float)H);
    if (band_top < 1) band_top = 1;
    if (band_top > H - 1) band_top = H - 1;
    const float r = synth_u01(h1);
    int u = (int)(h0 % (uint64_t)W);
    int v;
    const float t_road = c.road_frac, t_obj = t_road + c.object_frac, t_above = t_obj + c.above_band_frac;
    if (r < t_road) {
        const int top = (2 * H) / 3;  // lower third of the image
        v = top + (int)(h2 % (uint64_t)(H - top));
    } else if (r < t_obj && synth_object_feature(c, seed, frame, i, tables, u, v)) {
        // u, v set
    } else if (r >= t_obj && r < t_above) {
        v = (int)(h2 % (uint64_t)band_top);
    } else {
        v = band_top + (int)(h2 % (uint64_t)(H - band_top));
    }
    out[0] = (double)u;
    out[1] = (double)v;
}
