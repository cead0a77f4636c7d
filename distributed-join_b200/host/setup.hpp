// setup.hpp -- reference interface src/setup.hpp:28-59: device selection, memory pool and
// communicator construction by name.
#pragma once

#include <cstdint>
#include <string>

#include "communicator.hpp"
#include "cudf_shim.hpp"

// UCX pre-registration has no role on NVLink; the type only keeps the drivers' signatures intact.
class registered_memory_resource : public rmm::mr::device_memory_resource {};

void set_cuda_device();

void setup_memory_pool_and_communicator(
  Communicator*& communicator, registered_memory_resource*& registered_mr,
  rmm::mr::pool_memory_resource<rmm::mr::device_memory_resource>*& pool_mr, std::string communicator_name,
  std::string registration_method, int64_t communicator_buffer_size);

void destroy_memory_pool_and_communicator(
  Communicator* communicator, registered_memory_resource* registered_mr,
  rmm::mr::pool_memory_resource<rmm::mr::device_memory_resource>* pool_mr, std::string communicator_name,
  std::string registration_method);
