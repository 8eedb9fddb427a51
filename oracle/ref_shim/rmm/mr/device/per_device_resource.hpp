#pragma once
#include <rmm/resource_ref.hpp>
namespace rmm { namespace mr {
inline device_async_resource_ref get_current_device_resource() { return {}; }
inline device_async_resource_ref get_current_device_resource_ref() { return {}; }
}}  // namespace rmm::mr
