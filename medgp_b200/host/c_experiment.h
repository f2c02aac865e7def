// c_experiment.h -- experiment description: exp_setup.json, hyp_bound.txt, patient files,
// random initialisation and the output writers.  Same keys, file formats, RNG call order and
// method names as medgpc/src/dataio/c_experiment.{h,cpp}; the JSON reader is a small built-in
// parser (the reference uses rapidjson, which is not vendored).
#ifndef MEDGP_C_EXPERIMENT_H
#define MEDGP_C_EXPERIMENT_H

#include <string>
#include <vector>

class c_experiment {
  public:
    c_experiment();
    explicit c_experiment(const std::string &input_cfg_name);
    void init_default_param();

    std::string get_exp_train_dir() const { return exp_train_dir; }
    std::string get_exp_test_dir() const { return exp_test_dir; }
    std::string get_exp_data_dir() const { return exp_data_dir; }
    int get_kernel_index() const { return kernel_index; }
    std::vector<int> get_kernel_param() const { return kernel_param; }
    std::vector<int> get_feature_index() const { return feature_index; }
    std::vector<int> get_lik_param() const;
    int get_prior_mode() const { return prior_mode; }
    int get_cv_fold_num() const { return cv_fold_num; }
    int get_scg_init_num() const { return scg_init_num; }
    int get_scg_max_iter_num() const { return scg_max_iter_num; }
    int get_prior_sub_opt_iter() const { return prior_sub_opt_iter; }
    double get_online_learn_rate() const { return learn_rate; }
    double get_online_momentum() const { return momentum; }
    int get_random_seed() const { return srand_seed; }
    std::vector<float> get_prior_hyp() const { return prior_hyp; }

    void get_one_patient_data(std::string PAN, std::vector<int> &meta_vec, std::vector<float> &time_vec,
                              std::vector<float> &value_vec, bool verbose = true) const;

    // ---- cohort-level input: a cohort is thousands of patients x one small
    // text file per feature.  A shard first reads the SIZES of all patients (the count line of
    // every feature file) to deal patients to shards, then loads only its own patients; both
    // passes run over the host cores, and the per-feature statistics are read once.
    std::vector<int> get_cohort_sizes(const std::vector<std::string> &pans) const;
    struct patient_data {
        std::vector<int> meta;
        std::vector<float> time, value;
    };
    void get_cohort_data(const std::vector<std::string> &pans, std::vector<patient_data> &out) const;

    int get_hyp_num() const { return get_lik_num() + get_cov_num() + get_mean_num(); }
    int get_cov_num() const;
    int get_lik_num() const;
    int get_mean_num() const { return 0; }  // zero mean only (c_experiment.cpp:389-393)
    void get_hyp_bounds();
    void get_global_hyp(std::vector<std::vector<double> > &global_hyp_array);
    double get_one_random(const double &lb, const double &ub, const double &scale, const bool &flag_inv,
                          const bool &flag_log);
    void get_hyp_LMC_SM(std::vector<double> &hyp_array);
    void print_experiment() const;

    void output_double_bin(const std::string &file_prefix, std::vector<double> hyp_array) const;
    void output_int_txt(const std::string &file_prefix, std::vector<int> int_array) const;

    std::vector<int> get_test_kernel_param(int fold, const std::string &kernel_clust_alg) const;
    int get_test_cov_num(int fold, const std::string &kernel_clust_alg) const;
    std::vector<double> get_test_mode_param(int fold, const std::string &kernel_clust_alg) const;

  private:
    std::string exp_cfg_file, exp_data_dir, exp_hyp_bound_file, exp_top_dir, exp_train_dir, exp_test_dir,
        exp_kernel_dir;
    int srand_seed, kernel_index;
    std::vector<int> kernel_param, feature_index;
    int cv_fold_num, scg_init_num, scg_max_iter_num, prior_mode, prior_sub_opt_iter;
    double learn_rate, momentum;
    std::vector<float> prior_hyp;
    std::vector<double> hyp_array_ub, hyp_array_lb;
};

// Longest-processing-time-first deal of patients to shards, load ~ n^3:
// returns the shard of every patient.  Same rule as medgp_b200/shard.py: lpt_assign (ties by
// original order, first least-loaded shard).
std::vector<int> medgp_lpt_assign(const std::vector<int> &sizes, int nshard);

#endif
