// Multi-GPU halo exchange (see yb_halo.h).  Round-1 first cut: single rank only.
#include "yb_halo.h"

namespace yb {

struct HaloState {};

void halo_free(HaloState* h) { delete h; }
int halo_prepare(Solution&) { return set_error(YB_EUNSUPPORTED, "multi-rank runs are not implemented yet"); }
void halo_mark_dirty(Solution&, int) {}
int halo_exchange_all(Solution&, cudaStream_t) { return 0; }
int halo_run_stage(Solution&, int, int64_t, cudaStream_t) { return set_error(YB_EUNSUPPORTED, "multi-rank runs are not implemented yet"); }

}  // namespace yb

extern "C" {
int yb_halo_export_size(const yb_solution*, size_t* n) { if (n) *n = 0; return yb::set_error(YB_EUNSUPPORTED, "not implemented"); }
int yb_halo_export(yb_solution*, void*, size_t) { return yb::set_error(YB_EUNSUPPORTED, "not implemented"); }
int yb_halo_import(yb_solution*, int64_t, const void*, size_t) { return yb::set_error(YB_EUNSUPPORTED, "not implemented"); }
int yb_halo_finalize(yb_solution*) { return yb::set_error(YB_EUNSUPPORTED, "not implemented"); }
int yb_exchange_halos(yb_solution*) { return 0; }
}
