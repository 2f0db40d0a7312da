// Drop-in check with the reference's OWN caller: tracklets_depth::TrackletDepthModule (tracklets_depth/src/
// tracklet_depth_module.cpp + include/tracklets_depth/tracklet_depth_module.h, compiled UNMODIFIED from /root/reference against
// shim/include, see shim/Makefile target caller_dropin) drives Mono_Lidar::DepthEstimator exactly as the ROS node does:
// process(cloud, tracklets, camera info, semantic image) per frame -- SemanticPlane per frame, CalculateDepth for the previous
// and the current cloud -- and the depths it stores in its tracklets are compared with direct calls into the shim.
// ROS / OpenCV / feature_tracking types come from the stand-ins in tests/stubs_ros.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <tracklets_depth/tracklet_depth_module.h>

#include "mld_synth.h"

using Cloud = pcl::PointCloud<pcl::PointXYZI>;

static Cloud::Ptr make_cloud(const mld_synth_config& cfg, int frame) {
    auto cloud = std::make_shared<Cloud>();
    cloud->points.resize((size_t)mld_synth_points_per_frame(&cfg));
    static_assert(sizeof(pcl::PointXYZI) == 32, "PointXYZI layout");
    if (mld_synth_points_host_xyzi32(&cfg, 4711, frame, reinterpret_cast<float*>(cloud->points.data())) != 0) std::abort();
    return cloud;
}

int main() {
    mld_synth_config cfg;
    mld_synth_config_for(&cfg, 0, 1);  // KITTI shape, road / non-road feature mix
    const int W = cfg.image_width, H = cfg.image_height, F = 600;

    Mono_Lidar::DepthEstimatorParameters params;  // struct defaults + the yaml's values that matter here
    params.pixelarea_search_witdh = 6;
    params.pixelarea_search_height = 9;
    params.radiusSearch_count_min = 1;
    params.histogram_segmentation_bin_witdh = 0.3;
    params.viewray_plane_orthoganality_treshold = 0.03;
    params.ransac_plane_refinement_treshold = 0.2;  // the caller hands this to SemanticPlane as the inlier threshold
    tracklets_depth::TrackletDepthModule module(params);
    Eigen::Affine3d T;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T.matrix()(r, c) = cfg.cam_T[r * 4 + c];
    module.SetCameraLidarTransform(T);
    module.InitDepthEstimatorPre();

    auto info = std::make_shared<sensor_msgs::CameraInfo>();
    info->width = (uint32_t)W;
    info->height = (uint32_t)H;
    info->K[0] = info->K[4] = cfg.cam_f;
    info->K[2] = cfg.cam_cx;
    info->K[5] = cfg.cam_cy;
    info->K[8] = 1;
    auto label = std::make_shared<sensor_msgs::Image>();
    label->width = (uint32_t)W;
    label->height = (uint32_t)H;
    label->step = (uint32_t)W;
    label->encoding = "mono8";
    label->data.assign((size_t)W * H, 0);
    for (int y = 230; y < H; y++)
        for (int x = 0; x < W; x++) label->data[(size_t)y * W + x] = 7;  // road

    // an independent estimator for the cross-check (same parameters, same inputs)
    Mono_Lidar::DepthEstimator direct;
    direct.InitConfig(std::make_shared<Mono_Lidar::DepthEstimatorParameters>(params));
    direct.Initialize(std::make_shared<CameraPinhole>(W, H, cfg.cam_f, cfg.cam_cx, cfg.cam_cy), T);

    std::vector<double> uv((size_t)F * 2);
    Cloud::Ptr prev_cloud;
    Mono_Lidar::GroundPlane::Ptr gp_prev;
    long long with_depth = 0, checked = 0, road = 0;
    for (int frame = 0; frame < 3; frame++) {
        Cloud::Ptr cloud = make_cloud(cfg, frame);
        // every frame: F brand-new tracklets, each with its newest feature (this frame) and one in the previous frame
        if (mld_synth_features_host(&cfg, 4711, frame, F, uv.data()) != 0) return 2;
        auto msg = std::make_shared<matches_msg_ros::MatchesMsg>();
        msg->header.stamp = ros::Time(100 + frame, 0);
        for (int i = 0; i < F; i++) {
            matches_msg_ros::Tracklet t;
            t.id = (uint64_t)(frame * 100000 + i);
            matches_msg_ros::FeaturePoint cur, old;
            cur.u = (float)uv[(size_t)i * 2];
            cur.v = (float)uv[(size_t)i * 2 + 1];
            old.u = cur.u + 1.f < (float)W ? cur.u + 1.f : cur.u;  // where the feature was one frame ago
            old.v = cur.v;
            t.feature_points = {cur, old};
            msg->tracks.push_back(t);
        }
        module.process(cloud, msg, info, label);

        // what the module stored: tracklet i = {current feature (depth from `cloud`), previous feature (depth from prev_cloud)}
        std::vector<u_int64_t> ids;
        for (int i = 0; i < F; i++) ids.push_back((u_int64_t)(frame * 100000 + i));
        matches_msg_depth_ros::MatchesMsg out;
        module.convert_tracklets_to_matches_msg(msg, ids, out);
        if ((int)out.tracks.size() != F) return 3;

        // the same two CalculateDepth calls made directly (the reference's sequence, tracklet_depth_module.cpp:318, :330)
        Eigen::Matrix2Xd fc(2, F), fl(2, F);
        for (int i = 0; i < F; i++) {
            fc(0, i) = (int)msg->tracks[(size_t)i].feature_points[0].u;
            fc(1, i) = (int)msg->tracks[(size_t)i].feature_points[0].v;
            fl(0, i) = (int)msg->tracks[(size_t)i].feature_points[1].u;
            fl(1, i) = (int)msg->tracks[(size_t)i].feature_points[1].v;
        }
        Mono_Lidar::SemanticPlane::Camera cam;
        cam.f = cfg.cam_f;
        cam.cu = cfg.cam_cx;
        cam.cv = cfg.cam_cy;
        cam.transform_cam_lidar = T;
        cv::Mat img(H, W, CV_8UC1, label->data.data());
        Eigen::VectorXd d_cur, d_last;
        Eigen::VectorXi s_cur;
        if (prev_cloud) {
            Cloud::ConstPtr pc = prev_cloud;
            direct.CalculateDepth(pc, fl, d_last, gp_prev);
        }
        Mono_Lidar::GroundPlane::Ptr gp =
            std::make_shared<Mono_Lidar::SemanticPlane>(img, cam, std::set<int>{6, 7, 8, 9}, params.ransac_plane_refinement_treshold);
        Cloud::ConstPtr cc = cloud;
        direct.CalculateDepth(cc, fc, d_cur, s_cur, gp);
        for (int i = 0; i < F; i++) {
            const auto& fp = out.tracks[(size_t)i].feature_points;  // push_front order: newest first
            if (fp.size() != 2) return 4;
            const float want_cur = (float)d_cur(i), want_last = prev_cloud ? (float)d_last(i) : -1.f;
            if (fp[0].d != want_cur || fp[1].d != want_last) {
                std::printf("frame %d feature %d: module (%g, %g) direct (%g, %g)\n", frame, i, fp[0].d, fp[1].d, want_cur, want_last);
                return 5;
            }
            if (fp[0].d >= 0) with_depth++;
            if (s_cur(i) == Mono_Lidar::SuccessRoad) road++;
            checked += 2;
        }
        auto stats = direct.getDepthCalcStats();
        if (stats.getPointCount() != F || stats.getSuccess() + stats.getSuccessRoad() <= 0) return 6;
        gp_prev = gp;
        prev_cloud = cloud;
    }
    // the inline getters of tracklet_depth_module.h:109-123 (what the ROS node's debug publishers call)
    Cloud::Ptr cam_cloud = std::make_shared<Cloud>(), interp = std::make_shared<Cloud>();
    Eigen::Matrix2Xd vis;
    module.getCloudCameraCs(cam_cloud);
    module.getCloudInterpolated(interp);
    module.getPointsCloudImageCs(vis);
    auto st = module.getDepthCalcStats();
    if (cam_cloud->points.size() != (size_t)mld_synth_points_per_frame(&cfg) || !interp->points.empty() || vis.cols() <= 1000) return 7;
    std::printf("caller drop-in ok: %lld depths checked, %lld features with depth, %lld SuccessRoad, %d visible points, stats point count %d\n",
                checked, with_depth, road, vis.cols(), st.getPointCount());
    return 0;
}
