// Internal declarations shared by the translation units of libhft_sm100.so.
#pragma once
#include "../../include/hft_sm100.h"

namespace hft {
void count_launch(int n = 1);
void reset_launch_count();
// Brackets one kernel launch with events when profiling is enabled (hft_profile_enable); always counts the launch.
struct LaunchScope {
  int kclass;
  void* stream;
  void* ev0;
  LaunchScope(int kclass, void* stream);
  ~LaunchScope();
};
}  // namespace hft
