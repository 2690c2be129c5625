// Stand-in for include/config.h: the accessors the shim reads, same names and signatures (static, by reference).
#pragma once
class Config {
public:
    static Config& getInstance() { static Config c; return c; }
    static bool& hasPoints() { return getInstance().has_points; }
    static bool& hasLines() { return getInstance().has_lines; }
    static bool& bestLRMatches() { return getInstance().best_lr_matches; }
    static bool& useMotionModel() { return getInstance().use_motion_model; }
    static int& minFeatures() { return getInstance().min_features; }
    static int& maxIters() { return getInstance().max_iters; }
    static int& maxItersRef() { return getInstance().max_iters_ref; }
    static double& minRatio12P() { return getInstance().min_ratio_12_p; }
    static double& minRatio12L() { return getInstance().min_ratio_12_l; }
    static double& homogTh() { return getInstance().homog_th; }
    static double& minError() { return getInstance().min_error; }
    static double& minErrorChange() { return getInstance().min_error_change; }
    static double& inlierK() { return getInstance().inlier_k; }
    static double& lsdScale() { return getInstance().lsd_scale; }
    static double& orbScaleFactor() { return getInstance().orb_scale_factor; }
    bool has_points = true, has_lines = true, best_lr_matches = true, use_motion_model = false;
    int min_features = 10, max_iters = 5, max_iters_ref = 10;
    double min_ratio_12_p = 0.9, min_ratio_12_l = 0.9, homog_th = 1e-7, min_error = 1e-7, min_error_change = 1e-7, inlier_k = 4.0;
    double lsd_scale = 1.2, orb_scale_factor = 1.2;
};
