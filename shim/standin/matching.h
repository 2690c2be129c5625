// Stand-in for include/matching.h:36-60 (the two functions the shim defines).
#pragma once
#include <vector>
#include <opencv2/core.hpp>
#include "config.h"
namespace StVO {
int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
}  // namespace StVO
