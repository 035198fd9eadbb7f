// Stand-in for <boost/algorithm/string.hpp> — TEST INFRASTRUCTURE ONLY.  The reference's
// src/base/camera_models.cc includes this header but calls nothing from it.
#pragma once
