// Host build of efficientteacher_b200/csrc/loss_math.h for the CPU unit test of the hand-written CIoU backward.
// Test infrastructure only (never linked into libetb200.so).
#include "../../efficientteacher_b200/csrc/loss_math.h"
extern "C" void hm_row_ciou(const float* logits, const float* anch, const float* tbox, int n, float* ciou, float* grad) {
  for (int i = 0; i < n; ++i) ciou[i] = etb_row_ciou(logits + 4 * i, anch[2 * i], anch[2 * i + 1], tbox + 4 * i, grad + 4 * i);
}
extern "C" float hm_bce(float x, float z) { return etb_bce_logits(x, z); }
