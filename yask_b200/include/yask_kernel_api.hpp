// yask_kernel_api.hpp -- C++ host mirror of the reference's kernel API for the B200 engine.
//
// Same namespace, class names, method names, argument meaning and error behaviour as the reference's
// public kernel API (/root/reference/include/yask_kernel_api.hpp, yask_common_api.hpp,
// aux/yk_solution_api.hpp, aux/yk_var_api.hpp), so that user code written against the reference compiles
// unchanged against this header and links with libyask_kernel.<stencil>.b200.so.  This file was
// written from that interface (declaration order kept), not copied from it; the documentation of
// every call is the reference's.  The implementation (yask_b200/csrc/yk_api.cpp) is a thin adapter
// over the C ABI in include/yask_b200.h.
#pragma once

#include <cstdint>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#ifndef MPI_VERSION
typedef int MPI_Comm;   // as the reference does when MPI is not in use (yask_kernel_api.hpp:38-40)
#endif

namespace yask {

// ---- common API (yask_common_api.hpp) ------------------------------------------------------------------
typedef std::int64_t idx_t;
typedef std::vector<idx_t> idx_t_vec;
typedef std::initializer_list<idx_t> idx_t_init_list;
typedef std::vector<std::string> string_vec;

std::string yask_get_version_string();

class yask_output;
class yask_file_output;
class yask_string_output;
class yask_stdout_output;
class yask_null_output;
typedef std::shared_ptr<yask_output> yask_output_ptr;
typedef std::shared_ptr<yask_file_output> yask_file_output_ptr;
typedef std::shared_ptr<yask_string_output> yask_string_output_ptr;
typedef std::shared_ptr<yask_stdout_output> yask_stdout_output_ptr;
typedef std::shared_ptr<yask_null_output> yask_null_output_ptr;

class yask_exception : public std::exception {
    std::string _msg;
  public:
    yask_exception() : _msg("YASK exception") {}
    yask_exception(const std::string& message) : _msg(message) {}
    virtual ~yask_exception() {}
    virtual const char* what() const noexcept;
    virtual void add_message(const std::string& message);
    virtual const char* get_message() const;
};

#define THROW_YASK_EXCEPTION(message) do { yask::yask_exception e_(message); throw e_; } while (0)
#define FORMAT_AND_THROW_YASK_EXCEPTION(message) do { std::stringstream err_; err_ << message; THROW_YASK_EXCEPTION(err_.str()); } while (0)

class yask_output_factory {
  public:
    virtual ~yask_output_factory() {}
    virtual yask_file_output_ptr new_file_output(const std::string& file_name) const;
    virtual yask_string_output_ptr new_string_output() const;
    virtual yask_stdout_output_ptr new_stdout_output() const;
    virtual yask_null_output_ptr new_null_output() const;
};
class yask_output {
  public:
    virtual ~yask_output() {}
    virtual std::ostream& get_ostream() = 0;
};
class yask_file_output : public virtual yask_output {
  public:
    virtual ~yask_file_output() {}
    virtual std::string get_filename() const = 0;
    virtual void close() = 0;
};
class yask_string_output : public virtual yask_output {
  public:
    virtual ~yask_string_output() {}
    virtual std::string get_string() const = 0;
    virtual void discard() = 0;
};
class yask_stdout_output : public virtual yask_output {
  public:
    virtual ~yask_stdout_output() {}
};
class yask_null_output : public virtual yask_output {
  public:
    virtual ~yask_null_output() {}
};

void yask_print_splash(std::ostream& os, int argc, char** argv, std::string invocation_leader = "invocation: ");

// ---- kernel API -------------------------------------------------------------------------------------------
class yk_env;
class yk_solution;
class yk_var;
class yk_stats;
typedef std::shared_ptr<yk_env> yk_env_ptr;
typedef std::shared_ptr<yk_solution> yk_solution_ptr;
typedef std::shared_ptr<yk_var> yk_var_ptr;
typedef std::shared_ptr<yk_stats> yk_stats_ptr;

const int yask_numa_local = -1;
const int yask_numa_interleave = -2;
const int yask_numa_none = -9;
const int yask_numa_offload = -11;

class yk_factory {
  public:
    yk_factory();
    virtual ~yk_factory() {}
    virtual std::string get_version_string();
    virtual yk_env_ptr new_env() const;
    virtual yk_env_ptr new_env(MPI_Comm comm) const;
    virtual yk_solution_ptr new_solution(yk_env_ptr env) const;
    virtual yk_solution_ptr new_solution(yk_env_ptr env, const yk_solution_ptr source) const;
};

class yk_env {
  public:
    virtual ~yk_env() {}
    static void set_debug_output(yask_output_ptr debug);
    static void disable_debug_output();
    static yask_output_ptr get_debug_output();
    static inline void print_splash(int argc, char** argv, std::string invocation_leader = "invocation: ") {
        yask_print_splash(get_debug_output()->get_ostream(), argc, argv, invocation_leader);
    }
    static void set_trace_enabled(bool enable);
    static bool is_trace_enabled();
    virtual int get_num_ranks() const = 0;
    virtual int get_rank_index() const = 0;
    virtual void global_barrier() const = 0;
    virtual idx_t sum_over_ranks(idx_t rank_val) const = 0;
    virtual void assert_equality_over_ranks(idx_t rank_val, const std::string& descr) const = 0;
    virtual void finalize() = 0;
    [[noreturn]] virtual void exit(int code) = 0;
};

class yk_solution {
  public:
    virtual ~yk_solution() {}
    virtual const std::string& get_name() const = 0;
    virtual const std::string& get_description() const = 0;
    virtual std::string get_target() const = 0;
    virtual bool is_offloaded() const = 0;
    virtual int get_element_bytes() const = 0;
    virtual std::string get_step_dim_name() const = 0;
    virtual int get_num_domain_dims() const = 0;
    virtual string_vec get_domain_dim_names() const = 0;
    virtual string_vec get_misc_dim_names() const = 0;
    virtual void set_rank_domain_size(const std::string& dim, idx_t size) = 0;
    virtual void set_rank_domain_size_vec(const idx_t_vec& vals) = 0;
    virtual void set_rank_domain_size_vec(const idx_t_init_list& vals) = 0;
    virtual idx_t get_rank_domain_size(const std::string& dim) const = 0;
    virtual idx_t_vec get_rank_domain_size_vec() const = 0;
    virtual void set_overall_domain_size(const std::string& dim, idx_t size) = 0;
    virtual void set_overall_domain_size_vec(const idx_t_vec& vals) = 0;
    virtual void set_overall_domain_size_vec(const idx_t_init_list& vals) = 0;
    virtual idx_t get_overall_domain_size(const std::string& dim) const = 0;
    virtual idx_t_vec get_overall_domain_size_vec() const = 0;
    virtual void set_block_size(const std::string& dim, idx_t size) = 0;
    virtual void set_block_size_vec(const idx_t_vec& vals) = 0;
    virtual void set_block_size_vec(const idx_t_init_list& vals) = 0;
    virtual idx_t get_block_size(const std::string& dim) const = 0;
    virtual idx_t_vec get_block_size_vec() const = 0;
    virtual void set_num_ranks(const std::string& dim, idx_t num) = 0;
    virtual void set_num_ranks_vec(const idx_t_vec& vals) = 0;
    virtual void set_num_ranks_vec(const idx_t_init_list& vals) = 0;
    virtual idx_t get_num_ranks(const std::string& dim) const = 0;
    virtual idx_t_vec get_num_ranks_vec() const = 0;
    virtual void set_rank_index(const std::string& dim, idx_t num) = 0;
    virtual void set_rank_index_vec(const idx_t_vec& vals) = 0;
    virtual void set_rank_index_vec(const idx_t_init_list& vals) = 0;
    virtual idx_t get_rank_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_rank_index_vec() const = 0;
    virtual int get_num_outer_threads() const = 0;
    virtual int get_num_inner_threads() const = 0;
    virtual std::string apply_command_line_options(const std::string& args) = 0;
    virtual std::string apply_command_line_options(int argc, char* argv[]) = 0;
    virtual std::string apply_command_line_options(const string_vec& args) = 0;
    virtual std::string get_command_line_help() = 0;
    virtual std::string get_command_line_values() = 0;
    virtual int get_num_vars() const = 0;
    virtual yk_var_ptr get_var(const std::string& name) = 0;
    virtual std::vector<yk_var_ptr> get_vars() = 0;
    virtual void prepare_solution() = 0;
    virtual idx_t get_first_rank_domain_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_first_rank_domain_index_vec() const = 0;
    virtual idx_t get_last_rank_domain_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_last_rank_domain_index_vec() const = 0;
    virtual void run_solution(idx_t first_step_index, idx_t last_step_index) = 0;
    virtual void run_solution(idx_t step_index) = 0;
    virtual void copy_vars_to_device() const = 0;
    virtual void copy_vars_from_device() const = 0;
    virtual void exchange_halos() = 0;
    virtual void end_solution() = 0;
    virtual yk_stats_ptr get_stats() = 0;
    virtual void clear_stats() = 0;
    virtual void reset_auto_tuner(bool enable, bool verbose = false) = 0;
    virtual bool is_auto_tuner_enabled() const = 0;
    virtual void run_auto_tuner_now(bool verbose = true) = 0;
    virtual void set_min_pad_size(const std::string& dim, idx_t size) = 0;
    virtual idx_t get_min_pad_size(const std::string& dim) const = 0;
    virtual yk_var_ptr new_var(const std::string& name, const string_vec& dims) = 0;
    virtual yk_var_ptr new_var(const std::string& name, const std::initializer_list<std::string>& dims) = 0;
    virtual yk_var_ptr new_fixed_size_var(const std::string& name, const string_vec& dims, const idx_t_vec& dim_sizes) = 0;
    virtual yk_var_ptr new_fixed_size_var(const std::string& name, const std::initializer_list<std::string>& dims,
                                          const idx_t_init_list& dim_sizes) = 0;
    virtual bool set_default_numa_preferred(int numa_node) = 0;
    virtual int get_default_numa_preferred() const = 0;
    typedef std::function<void(yk_solution& soln)> hook_fn_t;
    typedef std::function<void(yk_solution& soln, idx_t first_step_index, idx_t last_step_index)> hook_fn_2idx_t;
    virtual void call_before_prepare_solution(hook_fn_t hook_fn) = 0;
    virtual void call_after_prepare_solution(hook_fn_t hook_fn) = 0;
    virtual void call_before_run_solution(hook_fn_2idx_t hook_fn) = 0;
    virtual void call_after_run_solution(hook_fn_2idx_t hook_fn) = 0;
    virtual void fuse_vars(yk_solution_ptr source) = 0;
    virtual void set_step_wrap(bool do_wrap) = 0;
    virtual bool get_step_wrap() const = 0;
    virtual void set_debug_output(yask_output_ptr debug) = 0;
};

class yk_stats {
  public:
    virtual ~yk_stats() {}
    virtual idx_t get_num_elements() = 0;
    virtual idx_t get_num_steps_done() = 0;
    virtual idx_t get_num_writes_done() = 0;
    virtual idx_t get_est_fp_ops_done() = 0;
    virtual double get_elapsed_secs() = 0;
};

class yk_var {
  public:
    virtual ~yk_var() {}
    virtual const std::string& get_name() const = 0;
    virtual int get_num_dims() const = 0;
    virtual string_vec get_dim_names() const = 0;
    virtual int get_num_domain_dims() const = 0;
    virtual bool is_dim_used(const std::string& dim) const = 0;
    virtual bool is_fixed_size() const = 0;
    virtual idx_t get_first_local_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_first_local_index_vec() const = 0;
    virtual idx_t get_last_local_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_last_local_index_vec() const = 0;
    virtual idx_t get_alloc_size(const std::string& dim) const = 0;
    virtual idx_t_vec get_alloc_size_vec() const = 0;
    virtual idx_t get_first_valid_step_index() const = 0;
    virtual idx_t get_last_valid_step_index() const = 0;
    virtual idx_t get_rank_domain_size(const std::string& dim) const = 0;
    virtual idx_t_vec get_rank_domain_size_vec() const = 0;
    virtual idx_t get_first_rank_domain_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_first_rank_domain_index_vec() const = 0;
    virtual idx_t get_last_rank_domain_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_last_rank_domain_index_vec() const = 0;
    virtual idx_t get_left_halo_size(const std::string& dim) const = 0;
    virtual idx_t get_right_halo_size(const std::string& dim) const = 0;
    virtual idx_t get_first_rank_halo_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_first_rank_halo_index_vec() const = 0;
    virtual idx_t get_last_rank_halo_index(const std::string& dim) const = 0;
    virtual idx_t_vec get_last_rank_halo_index_vec() const = 0;
    virtual idx_t get_left_pad_size(const std::string& dim) const = 0;
    virtual idx_t get_right_pad_size(const std::string& dim) const = 0;
    virtual idx_t get_left_extra_pad_size(const std::string& dim) const = 0;
    virtual idx_t get_right_extra_pad_size(const std::string& dim) const = 0;
    virtual idx_t get_first_misc_index(const std::string& dim) const = 0;
    virtual idx_t get_last_misc_index(const std::string& dim) const = 0;
    virtual bool are_indices_local(const idx_t_vec& indices) const = 0;
    virtual bool are_indices_local(const idx_t_init_list& indices) const = 0;
    virtual double get_element(const idx_t_vec& indices) const = 0;
    virtual double get_element(const idx_t_init_list& indices) const = 0;
    virtual idx_t set_element(double val, const idx_t_vec& indices, bool strict_indices = true) = 0;
    virtual idx_t set_element(double val, const idx_t_init_list& indices, bool strict_indices = true) = 0;
    virtual idx_t get_elements_in_slice(float* buffer_ptr, size_t buffer_size, const idx_t_vec& first_indices,
                                        const idx_t_vec& last_indices) const = 0;
    virtual idx_t get_elements_in_slice(double* buffer_ptr, size_t buffer_size, const idx_t_vec& first_indices,
                                        const idx_t_vec& last_indices) const = 0;
    virtual idx_t add_to_element(double val, const idx_t_vec& indices, bool strict_indices = true) = 0;
    virtual idx_t add_to_element(double val, const idx_t_init_list& indices, bool strict_indices = true) = 0;
    virtual void set_all_elements_same(double val) = 0;
    virtual idx_t set_elements_in_slice_same(double val, const idx_t_vec& first_indices, const idx_t_vec& last_indices,
                                             bool strict_indices = true) = 0;
    virtual idx_t set_elements_in_slice(const float* buffer_ptr, size_t buffer_size, const idx_t_vec& first_indices,
                                        const idx_t_vec& last_indices) = 0;
    virtual idx_t set_elements_in_slice(const double* buffer_ptr, size_t buffer_size, const idx_t_vec& first_indices,
                                        const idx_t_vec& last_indices) = 0;
    virtual idx_t set_elements_in_slice(const yk_var_ptr source, const idx_t_vec& first_source_indices,
                                        const idx_t_vec& first_target_indices, const idx_t_vec& last_target_indices) = 0;

    // Reductions (aux/yk_var_api.hpp:963-1100).
    enum yk_reduction_idx { yk_sum_reduction_idx, yk_sum_squares_reduction_idx, yk_product_reduction_idx, yk_max_reduction_idx,
                            yk_min_reduction_idx };
    static constexpr int yk_sum_reduction = 1 << yk_sum_reduction_idx, yk_sum_squares_reduction = 1 << yk_sum_squares_reduction_idx,
                         yk_product_reduction = 1 << yk_product_reduction_idx, yk_max_reduction = 1 << yk_max_reduction_idx,
                         yk_min_reduction = 1 << yk_min_reduction_idx;
    class yk_reduction_result {
      public:
        virtual ~yk_reduction_result() {}
        virtual int get_reduction_mask() const = 0;
        virtual idx_t get_num_elements_reduced() const = 0;
        virtual double get_sum() const = 0;
        virtual double get_sum_squares() const = 0;
        virtual double get_product() const = 0;
        virtual double get_max() const = 0;
        virtual double get_min() const = 0;
    };
    typedef std::shared_ptr<yk_reduction_result> yk_reduction_result_ptr;
    virtual yk_reduction_result_ptr reduce_elements_in_slice(int reduction_mask, const idx_t_vec& first_indices,
                                                             const idx_t_vec& last_indices, bool strict_indices = true) = 0;

    virtual std::string format_indices(const idx_t_vec& indices) const = 0;
    virtual std::string format_indices(const idx_t_init_list& indices) const = 0;
    virtual int get_halo_exchange_l1_norm() const = 0;
    virtual void set_halo_exchange_l1_norm(int norm) = 0;
    virtual bool is_dynamic_step_alloc() const = 0;
    virtual bool set_numa_preferred(int numa_node) = 0;
    virtual int get_numa_preferred() const = 0;
    virtual void set_left_min_pad_size(const std::string& dim, idx_t size) = 0;
    virtual void set_right_min_pad_size(const std::string& dim, idx_t size) = 0;
    virtual void set_min_pad_size(const std::string& dim, idx_t size) = 0;
    virtual void set_left_halo_size(const std::string& dim, idx_t size) = 0;
    virtual void set_right_halo_size(const std::string& dim, idx_t size) = 0;
    virtual void set_halo_size(const std::string& dim, idx_t size) = 0;
    virtual void set_alloc_size(const std::string& dim, idx_t size) = 0;
    virtual void set_first_misc_index(const std::string& dim, idx_t idx) = 0;
    virtual bool is_storage_allocated() const = 0;
    virtual idx_t get_num_storage_bytes() const = 0;
    virtual idx_t get_num_storage_elements() const = 0;
    virtual void alloc_storage() = 0;
    virtual void release_storage() = 0;
    virtual bool is_storage_layout_identical(const yk_var_ptr other) const = 0;
    virtual void fuse_vars(yk_var_ptr source) = 0;
    virtual void* get_raw_storage_buffer() = 0;
    virtual idx_t get_elements_in_slice(void* buffer_ptr, const idx_t_vec& first_indices, const idx_t_vec& last_indices) const = 0;
    virtual idx_t set_elements_in_slice(const void* buffer_ptr, const idx_t_vec& first_indices, const idx_t_vec& last_indices) = 0;
    virtual idx_t get_first_rank_alloc_index(const std::string& dim) const { return get_first_local_index(dim); }
    virtual idx_t get_last_rank_alloc_index(const std::string& dim) const { return get_last_local_index(dim); }
};

}  // namespace yask
