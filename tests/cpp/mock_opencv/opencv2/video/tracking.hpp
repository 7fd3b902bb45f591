// mock: the tracking declarations the adapter needs live in core.hpp (TEST SCAFFOLDING ONLY)
#include <opencv2/core.hpp>
