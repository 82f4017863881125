// Forwarding header: the reference splits its API over myyuv_bmp.hpp and myyuv_yuv.hpp; here both classes live in myyuv.hpp.
#pragma once
#include "myyuv.hpp"
