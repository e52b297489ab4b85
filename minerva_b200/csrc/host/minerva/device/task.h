// Task / TaskData / PhysicalData / DeviceListener -- the contract between the scheduler and a device
// (reference: minerva/device/task.h, task_data.h, device_listener.h, op/physical.h:9-21), unchanged in shape.
#pragma once
#include <cstdint>
#include <vector>
#include "op/hotpath.h"

namespace minerva {

struct PhysicalData {
  PhysicalData(const Scale& s, uint64_t d, uint64_t id) : size(s), device_id(d), data_id(id) {}
  Scale size;
  uint64_t device_id;
  uint64_t data_id;
};

struct TaskData {
  TaskData(const PhysicalData& p, uint64_t i) : physical_data(p), id(i) {}
  PhysicalData physical_data;
  uint64_t id;   // only meaningful to the issuer of the task
};

struct Task {
  std::vector<TaskData> inputs;
  std::vector<TaskData> outputs;
  PhysicalOp op;
  uint64_t id = 0;
  bool light = false;
};

class DeviceListener {
 public:
  virtual ~DeviceListener() = default;
  virtual void OnOperationComplete(Task*) = 0;
};

}  // namespace minerva
