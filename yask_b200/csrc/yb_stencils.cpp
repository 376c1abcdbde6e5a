// Registry of built-in solutions: the engine-side counterpart of the reference's REGISTER_SOLUTION
// table (/root/reference/src/compiler/compiler_main.cpp:157-234 looks solutions up by "-stencil").
#include "yb_core.h"

namespace yb {

// emitter-generated solutions (yb_gen.cu)
int gen_registry_size();
const char* gen_registry_name(int i);
int gen_registry_create(const std::string& name, int elem_bytes, StencilSpec& spec, std::unique_ptr<Engine>& eng);

namespace {
struct Entry { const char* name; int default_radius; };
const Entry kEntries[] = {
    {"iso3dfd", 8},
};
}  // namespace

const int kNumBuiltin = int(sizeof(kEntries) / sizeof(kEntries[0]));
int registry_size() { return kNumBuiltin + gen_registry_size(); }
const char* registry_name(int i) {
    if (i < 0 || i >= registry_size()) return "";
    return i < kNumBuiltin ? kEntries[i].name : gen_registry_name(i - kNumBuiltin);
}

int registry_create(const std::string& name, int radius, int elem_bytes, StencilSpec& spec, std::unique_ptr<Engine>& eng) {
    if (name == "iso3dfd") {
        if (radius <= 0) radius = 8;
        if (elem_bytes == 0) elem_bytes = 4;
        if (radius > 8) return set_error(YB_EUNSUPPORTED, "iso3dfd: radius %d > 8 is not supported", radius);
        if (elem_bytes == 8) {
            // double precision: served by emitter-generated sweep kernels, radius 8 and radius 3 (the radius of the
            // reference's own fp64 validation run, /root/reference/src/kernel/Makefile:1155); the hand-written kernel is fp32
            if (radius != 8 && radius != 3) return set_error(YB_EUNSUPPORTED, "iso3dfd fp64 is generated for radius 8 and radius 3 only");
            int rc = gen_registry_create(radius == 8 ? "iso3dfd_fp64" : "iso3dfd_fp64_r3", 8, spec, eng);
            if (rc == 0) { spec.name = "iso3dfd"; spec.radius = radius; }
            return rc;
        }
        spec = iso3dfd_spec(radius, elem_bytes, false);
        eng = make_iso3dfd_engine();
        return 0;
    }
    if (name == "iso3dfd_sponge" && elem_bytes == 8) {
        // fp64: generated for radius 6, the reference's fp64 validation run of this solution (src/kernel/Makefile:1156)
        if (radius > 0 && radius != 6) return set_error(YB_EUNSUPPORTED, "iso3dfd_sponge fp64 is generated for radius 6 only");
        int rc = gen_registry_create("iso3dfd_sponge_fp64_r6", 8, spec, eng);
        if (rc == 0) { spec.name = "iso3dfd_sponge"; spec.radius = 6; }
        return rc;
    }
    int rc = gen_registry_create(name, elem_bytes, spec, eng);
    if (rc != YB_EINVAL) return rc;
    return set_error(YB_EINVAL, "unknown stencil solution '%s'", name.c_str());
}

}  // namespace yb
