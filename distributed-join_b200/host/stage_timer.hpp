// stage_timer.hpp -- wall-clock stage timing behind `report_timing`, printing the reference's
// scrapeable lines ("Rank r: <stage> takes X ms") from one place.
#pragma once

#include <iostream>
#include <string>

#include "bootstrap.hpp"

class StageTimer {
 public:
  StageTimer(bool enabled, int rank) : enabled_(enabled), rank_(rank), t0_(enabled ? dj_bootstrap::wtime() : 0.0) {}
  // prints the time since construction / the previous lap and restarts the clock
  void lap(const std::string& stage)
  {
    if (!enabled_) return;
    const double now = dj_bootstrap::wtime();
    std::cout << "Rank " << rank_ << ": " << stage << " takes " << (now - t0_) * 1e3 << "ms" << std::endl;
    t0_ = now;
  }

 private:
  bool enabled_;
  int rank_;
  double t0_;
};
