#pragma once
#include <rmm/cuda_stream_view.hpp>
