// mlctl.h -- how many quantizers the pruned scan's lower bound sums (plain C++, no CUDA: unit-tested on
// the CPU by tests/test_mlctl.py).
//
// Fewer quantizers make the bound pass cheaper and the survivor evaluation dearer; where the sum is
// smallest depends on the data, the range and the thresholds, so it is measured.  Every main-stage
// launch is timed; its cost (ms per 1e9 (row, query) pairs) is fed to `on_measurement`, and the size
// moves one step in the direction that last helped: an eighth of the current size, a quarter once a probe
// has paid off and until one fails (the failed quarter step is retried as an eighth from the best size
// before the other side is tried).  Two failed probes in a row park the controller at the best size for
// 24 launches (then 48, 96, ... up to 4096), and once it has parked, probes move by a single quantizer.  Only launches of one shape (tiles, rows) are compared;
// after 8 consecutive measurements of another shape the search restarts from the current size.  A cold
// controller starts with every quantizer (the bound of round 1b).
#pragma once
#include <algorithm>

namespace gulon {

struct MlController {
  int M = 0;                 // quantizers of the codebook
  int hint = 0;              // size the next launch uses
  int best = 0;              // best size measured so far (0: none yet) and its cost
  double best_cost = 0;
  int dir = -1;              // direction of the next probe
  int reversals = 0;         // failed probes since the last success
  int hold = 0;              // launches left before the next probe (parked while > 0)
  int hold_len = 24;
  bool coarse = false;       // a probe has paid off and none has failed yet: quarter steps
  bool failed = false;       // a probe has failed since the search (re)started
  bool fine = false;         // parked once: probes move by one quantizer
  long long shape = 0;       // shape key of the launches being compared
  int shape_miss = 0;

  int lo() const { return std::min(M, 4); }
  int step_from(int from, int d) const {
    const int stp = fine ? 1 : std::max(1, from / (coarse ? 4 : 8));
    return std::max(lo(), std::min(M, from + d * stp));
  }
  void restart_search() {
    best = 0;
    best_cost = 0;
    reversals = 0;
    hold = 0;
    hold_len = 24;
    fine = false;
    coarse = false;
    failed = false;
  }
  // (re)initialise for a codebook of M quantizers
  void reset(int M_) {
    M = M_;
    hint = M_;
    dir = -1;
    shape = 0;
    shape_miss = 0;
    restart_search();
  }
  bool searching() const { return hold == 0; }

  // One timed main-stage launch: `ml` quantizers, `cost` ms per 1e9 pairs, `shape_key` != 0.
  void on_measurement(int ml, double cost, long long shape_key) {
    if (ml != hint || !(cost > 0)) return;  // a launch enqueued before the last decision
    if (shape != shape_key) {
      if (shape == 0 || ++shape_miss >= 8) {
        shape = shape_key;  // a new workload: start comparing afresh from where we are
        shape_miss = 0;
        restart_search();
      } else {
        return;
      }
    } else {
      shape_miss = 0;
    }
    if (hold > 0) {
      best_cost = 0.75 * best_cost + 0.25 * cost;  // parked at the best size: keep its cost current
      return;
    }
    if (best == 0) {
      best = hint;
      best_cost = cost;
      hint = step_from(best, dir);
    } else if (hint != best && cost < best_cost * 0.985) {
      best = hint;  // the probe paid off: keep walking, faster while nothing has failed
      best_cost = cost;
      reversals = 0;
      hold_len = 24;
      if (!failed && !fine) coarse = true;
      hint = step_from(best, dir);
    } else if (hint == best) {
      best_cost = 0.5 * best_cost + 0.5 * cost;
      hint = step_from(best, dir);
    } else if (coarse) {
      coarse = false;  // a quarter step overshot: try an eighth from the best size, same direction
      failed = true;
      hint = step_from(best, dir);
    } else {
      dir = -dir;  // the probe did not pay off: try the other side of the best
      reversals++;
      failed = true;
      hint = step_from(best, dir);
    }
    if (hint == best) {
      // nowhere to go on this side (range limit): counts as a failed probe
      dir = -dir;
      reversals++;
      hint = step_from(best, dir);
    }
    if (reversals >= 2 || hint == best) {
      // both neighbours are worse: park, and for twice as long every time that happens again (a probe
      // next to a cliff -- unclustered data -- can cost several normal launches)
      reversals = 0;
      fine = true;
      hold = hold_len;
      hold_len = std::min(hold_len * 2, 4096);
      hint = best;
    }
  }

  // Called once per main-stage launch, after any measurement has been fed: the size to use.
  int next_launch() {
    if (hold > 0 && --hold == 0 && best > 0) {
      int probe = step_from(best, dir);
      if (probe == best) {
        dir = -dir;
        probe = step_from(best, dir);
      }
      hint = probe;  // (== best only when M leaves no room at all: the next measurement parks again)
    }
    return hint;
  }
};

}  // namespace gulon
