// TEST INFRASTRUCTURE ONLY -- never linked into, or called from, the product path.
//
// ref_driver: drives an UNMODIFIED reference YASK kernel library
// (libyask_kernel.<stencil>.<arch>.so, built out-of-tree by oracle/build_ref.sh from the
// sources where they lie under /root/reference) through the reference's public
// yk_* API only (include/yask_kernel_api.hpp).  It fills every var from raw binary
// files holding *logical-index* data (so the inputs do not depend on the reference's
// folded storage layout), runs run_solution(0, steps-1)
// (the optimized vector path: src/kernel/lib/context.cpp:220) and dumps the domain
// points of every var back to raw files.  tests/golden/make_golden.py uses it to
// produce the committed golden fixtures that pin oracle/*.c.
//
// usage:
//   ref_driver info  NX NY NZ            -> prints manifest (var geometry) on stdout
//   ref_driver run   NX NY NZ STEPS DIR  -> reads DIR/<var>.t<step>.in, writes DIR/<var>.t<step>.out
//                                           (vars without a step dim use ".t0")
#include "yask_kernel_api.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using namespace yask;

namespace {

struct VarGeom {
    std::string name;
    string_vec dims;
    bool has_step = false;
    idx_t first_step = 0, last_step = 0;
    // Per non-step dim, in var-declared order.
    idx_t_vec in_first, in_last;    // rank halo box (what the stencil may read).
    idx_t_vec out_first, out_last;  // rank domain box (what the stencil writes).
};

VarGeom geom_of(yk_solution_ptr soln, yk_var_ptr v) {
    VarGeom g;
    g.name = v->get_name();
    g.dims = v->get_dim_names();
    auto step_dim = soln->get_step_dim_name();
    auto ddims = soln->get_domain_dim_names();
    for (auto& d : g.dims) {
        if (d == step_dim) {
            g.has_step = true;
            g.first_step = v->get_first_valid_step_index();
            g.last_step = v->get_last_valid_step_index();
            continue;
        }
        bool is_domain = false;
        for (auto& dd : ddims)
            if (dd == d) is_domain = true;
        if (is_domain) {
            g.in_first.push_back(v->get_first_rank_halo_index(d));
            g.in_last.push_back(v->get_last_rank_halo_index(d));
            g.out_first.push_back(v->get_first_rank_domain_index(d));
            g.out_last.push_back(v->get_last_rank_domain_index(d));
        } else {
            g.in_first.push_back(v->get_first_misc_index(d));
            g.in_last.push_back(v->get_last_misc_index(d));
            g.out_first.push_back(v->get_first_misc_index(d));
            g.out_last.push_back(v->get_last_misc_index(d));
        }
    }
    return g;
}

size_t box_elems(const idx_t_vec& f, const idx_t_vec& l) {
    size_t n = 1;
    for (size_t i = 0; i < f.size(); i++) n *= size_t(l[i] - f[i] + 1);
    return n;
}

idx_t_vec with_step(const VarGeom& g, const idx_t_vec& box, idx_t t) {
    idx_t_vec r;
    size_t j = 0;
    for (size_t i = 0; i < g.dims.size(); i++) {
        if (g.has_step && i == 0) r.push_back(t);  // step dim is always first (yc API rule).
        else r.push_back(box[j++]);
    }
    return r;
}

template <typename T>
void load_var(yk_var_ptr v, const VarGeom& g, const std::string& dir) {
    idx_t t0 = g.has_step ? g.first_step : 0, t1 = g.has_step ? g.last_step : 0;
    size_t n = box_elems(g.in_first, g.in_last);
    std::vector<T> buf(n);
    for (idx_t t = t0; t <= t1; t++) {
        std::string fn = dir + "/" + g.name + ".t" + std::to_string(t) + ".in";
        std::ifstream f(fn, std::ios::binary);
        if (!f) { std::cerr << "ref_driver: cannot open " << fn << "\n"; exit(2); }
        f.read(reinterpret_cast<char*>(buf.data()), n * sizeof(T));
        if (size_t(f.gcount()) != n * sizeof(T)) { std::cerr << "ref_driver: short file " << fn << "\n"; exit(2); }
        if (g.dims.size() == (g.has_step ? 1u : 0u)) {
            // scalar var: no non-step dims.
            idx_t_vec idx;
            if (g.has_step) idx.push_back(t);
            v->set_element(double(buf[0]), idx);
        } else {
            v->set_elements_in_slice(buf.data(), n, with_step(g, g.in_first, t), with_step(g, g.in_last, t));
        }
    }
}

template <typename T>
void dump_var(yk_var_ptr v, const VarGeom& g0, yk_solution_ptr soln, const std::string& dir) {
    VarGeom g = geom_of(soln, v);  // valid step window may have moved.
    idx_t t0 = g.has_step ? g.first_step : 0, t1 = g.has_step ? g.last_step : 0;
    size_t n = box_elems(g.out_first, g.out_last);
    std::vector<T> buf(n);
    for (idx_t t = t0; t <= t1; t++) {
        if (g.dims.size() == (g.has_step ? 1u : 0u)) {
            idx_t_vec idx;
            if (g.has_step) idx.push_back(t);
            buf[0] = T(v->get_element(idx));
        } else {
            v->get_elements_in_slice(buf.data(), n, with_step(g, g.out_first, t), with_step(g, g.out_last, t));
        }
        std::string fn = dir + "/" + g.name + ".t" + std::to_string(t) + ".out";
        std::ofstream f(fn, std::ios::binary);
        f.write(reinterpret_cast<const char*>(buf.data()), n * sizeof(T));
    }
    (void)g0;
}

void print_manifest(yk_solution_ptr soln) {
    std::cout << "solution " << soln->get_name() << " elem_bytes " << soln->get_element_bytes()
              << " target " << soln->get_target() << "\n";
    for (auto v : soln->get_vars()) {
        auto g = geom_of(soln, v);
        std::cout << "var " << g.name << " ndims " << g.dims.size() << " dims";
        for (auto& d : g.dims) std::cout << " " << d;
        std::cout << " has_step " << int(g.has_step) << " steps " << g.first_step << " " << g.last_step;
        std::cout << " in_first";
        for (auto i : g.in_first) std::cout << " " << i;
        std::cout << " in_last";
        for (auto i : g.in_last) std::cout << " " << i;
        std::cout << " out_first";
        for (auto i : g.out_first) std::cout << " " << i;
        std::cout << " out_last";
        for (auto i : g.out_last) std::cout << " " << i;
        std::cout << "\n";
    }
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 5) {
        std::cerr << "usage: ref_driver info NX NY NZ | run NX NY NZ STEPS DIR\n";
        return 1;
    }
    std::string mode = argv[1];
    idx_t n[3] = {atoll(argv[2]), atoll(argv[3]), atoll(argv[4])};
    try {
        yk_factory kfac;
        auto env = kfac.new_env();
        yk_env::disable_debug_output();
        auto soln = kfac.new_solution(env);
        auto ddims = soln->get_domain_dim_names();
        for (size_t i = 0; i < ddims.size() && i < 3; i++) soln->set_overall_domain_size(ddims[i], n[i]);
        // Same no-tuner flags as the reference's own validation recipe (src/kernel/Makefile:1024-1028).
        soln->apply_command_line_options("-no-pre_auto_tune -no-auto_tune");
        soln->prepare_solution();

        if (mode == "info") {
            print_manifest(soln);
            soln->end_solution();
            return 0;
        }
        if (argc < 7) { std::cerr << "run needs STEPS DIR\n"; return 1; }
        idx_t steps = atoll(argv[5]);
        std::string dir = argv[6];
        bool f32 = soln->get_element_bytes() == 4;

        std::vector<VarGeom> geoms;
        for (auto v : soln->get_vars()) {
            auto g = geom_of(soln, v);
            geoms.push_back(g);
            if (f32) load_var<float>(v, g, dir);
            else load_var<double>(v, g, dir);
        }
        if (steps > 0) soln->run_solution(0, steps - 1);
        size_t k = 0;
        for (auto v : soln->get_vars()) {
            if (f32) dump_var<float>(v, geoms[k], soln, dir);
            else dump_var<double>(v, geoms[k], soln, dir);
            k++;
        }
        print_manifest(soln);
        soln->end_solution();
    } catch (yask_exception& e) {
        std::cerr << "ref_driver: YASK exception: " << e.get_message() << "\n";
        return 3;
    }
    return 0;
}
