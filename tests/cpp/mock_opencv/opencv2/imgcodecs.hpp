// mock: cv::ImreadModes lives in core.hpp (TEST SCAFFOLDING ONLY)
#include <opencv2/core.hpp>
