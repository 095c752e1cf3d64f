// rmm/cuda_stream_view.hpp stand-in — see cudf/detail/b2_bridge.hpp
#pragma once
#include "../cudf/detail/b2_bridge.hpp"
