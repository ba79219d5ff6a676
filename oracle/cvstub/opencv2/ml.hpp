// TEST INFRASTRUCTURE ONLY — stand-in for <opencv2/ml.hpp> (include/core.h:15 includes it; nothing on the path uses cv::ml).
#pragma once
#include "opencv.hpp"
