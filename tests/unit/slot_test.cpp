// TEST ONLY: step index -> storage slot map of a var with spare slots (yask_b200/csrc/yb_core.h, Var::slot_of / nslots / bytes).
// The API-visible behaviour must stay the reference's imod_flr(t, alloc_t) (/root/reference/src/kernel/lib/yk_var.hpp:131-147)
// whichever slot set is live.
#include <cstdio>
#include <cstdlib>

#include "../../yask_b200/csrc/yb_core.h"

#define CHECK(c) do { if (!(c)) { printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
    yb::Var v;
    yb::Dim t; t.spec.kind = yb::DIM_STEP; t.spec.name = "t";
    v.dims.push_back(t);
    v.spec.step_alloc = 2;
    v.slot_elems = 1000;
    v.elem_bytes = 4;
    CHECK(v.nslots() == 2 && v.bytes() == 2 * 1000 * 4);
    for (int s = -5; s <= 6; s++) CHECK(v.slot_of(s) == ((s % 2) + 2) % 2);
    v.extra_slots = 2;
    CHECK(v.step_alloc() == 2 && v.nslots() == 4 && v.bytes() == 4 * 1000 * 4);
    for (int bias = 0; bias <= 2; bias += 2) {
        v.slot_bias = bias;
        for (int s = -5; s <= 6; s++) {
            CHECK(v.slot_of(s) == ((s % 2) + 2) % 2 + bias);         // same wrap as the reference, inside the live pair
            CHECK(v.slot_of(s + 2) == v.slot_of(s));                  // steps two apart share storage (alloc_t = 2)
            CHECK((v.slot_of(s) ^ 1) == v.slot_of(s - 1));            // p(t-1) is the partner of p(t): what the tensor maps assume
            CHECK(((v.slot_of(s) ^ 2) & 2) != (bias & 2));            // ^2 addresses the spare pair
        }
    }
    // valid-step window is step_alloc long whatever the number of slots
    v.first_valid_step = 0;
    v.update_valid_step(7);
    CHECK(v.first_valid_step == 6 && v.last_valid_step() == 7);
    yb::Var w;       // a var without step dim
    w.extra_slots = 2;
    CHECK(w.nslots() == 1 && w.slot_of(5) == 0);
    printf("OK\n");
    return 0;
}
