// C wrappers around gulon::MlController (gulon_b200/csrc/mlctl.h) for tests/test_mlctl.py.
#include "mlctl.h"

extern "C" {
void *mlc_new(int M) {
  auto *c = new gulon::MlController();
  c->reset(M);
  return c;
}
void mlc_free(void *p) { delete static_cast<gulon::MlController *>(p); }
void mlc_measure(void *p, int ml, double cost, long long shape) {
  static_cast<gulon::MlController *>(p)->on_measurement(ml, cost, shape);
}
int mlc_next(void *p) { return static_cast<gulon::MlController *>(p)->next_launch(); }
int mlc_searching(void *p) { return static_cast<gulon::MlController *>(p)->searching() ? 1 : 0; }
int mlc_best(void *p) { return static_cast<gulon::MlController *>(p)->best; }
int mlc_hold(void *p) { return static_cast<gulon::MlController *>(p)->hold; }
}
