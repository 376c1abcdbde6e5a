// Registry of built-in solutions: the engine-side counterpart of the reference's REGISTER_SOLUTION
// table (/root/reference/src/compiler/compiler_main.cpp:157-234 looks solutions up by "-stencil").
#include "yb_core.h"

namespace yb {

namespace {
struct Entry { const char* name; int default_radius; };
const Entry kEntries[] = {
    {"iso3dfd", 8},
};
}  // namespace

int registry_size() { return int(sizeof(kEntries) / sizeof(kEntries[0])); }
const char* registry_name(int i) { return (i >= 0 && i < registry_size()) ? kEntries[i].name : ""; }

int registry_create(const std::string& name, int radius, int elem_bytes, StencilSpec& spec, std::unique_ptr<Engine>& eng) {
    if (name == "iso3dfd") {
        if (radius <= 0) radius = 8;
        if (radius > 8) return set_error(YB_EUNSUPPORTED, "iso3dfd: radius %d > 8 is not supported", radius);
        spec = iso3dfd_spec(radius, elem_bytes, false);
        eng = make_iso3dfd_engine();
        return 0;
    }
    return set_error(YB_EINVAL, "unknown stencil solution '%s'", name.c_str());
}

}  // namespace yb
