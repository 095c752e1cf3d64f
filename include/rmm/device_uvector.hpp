// rmm/device_uvector.hpp stand-in — see cudf/detail/b2_bridge.hpp
#pragma once
#include "../cudf/detail/b2_bridge.hpp"
