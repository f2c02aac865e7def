// main_one_test -- online one-step-ahead imputation of one patient, with and without online
// hyper-parameter updates (same CLI, semantics and output files as the reference's
// medgpc/src/main_one_test.cpp:45-480).
//   main_one_test --cfg exp_setup.json --pan <PAN> --thread <T> --fold <F> --kernclust-alg <A>
// Every held-out observation (tt, jj) trains on past + same-time observations and predicts
// one point.  Without updates theta is fixed and ONE factorisation of the time-ordered patient
// yields every prediction (medgp_cuda_predict_online); with updates theta moves at each update
// time, so the unit is one time stamp: its observations are predicted from one factorisation of
// "72 h history + this time stamp" (leave-one-out inside the time stamp).  Whenever that path
// does not apply, the training sets are refitted one by one as the reference does.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>

#include "c_experiment.h"
#include "imputation.h"
#include "medgp_host.h"

using std::cout;
using std::endl;
using std::string;
using std::vector;

#define CMD_NUM 11

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}


static void run_test_one(c_experiment &curr_exp, const string &PAN, int fold, bool flag_update,
                         const string &output_prefix, const string &kernel_clust_alg, c_kernel *&kptr,
                         c_meanfunc *&mptr, c_likelihood *&lptr, c_inference *&iptr, c_prior *&pptr)
{
    cout << "running online imputation: " << (flag_update ? "with" : "without") << " online updating" << endl
         << "testing patinet: " << PAN << " in cross-validation fold " << fold << endl;
    const vector<int> test_kernel_param = curr_exp.get_test_kernel_param(fold, kernel_clust_alg);
    vector<int> meta_array;
    vector<float> time_array, value_array;
    curr_exp.get_one_patient_data(PAN, meta_array, time_array, value_array);
    cout << "number of data points = " << time_array.size() << endl;
    const double t1 = now_s();
    bool test_flag = true;
    if (time_array.empty()) {
        cout << "Warning: no samples for testing" << endl;
        test_flag = false;
    } else {
        vector<float> unique_time_array(time_array);
        std::sort(unique_time_array.begin(), unique_time_array.end());
        unique_time_array.erase(std::unique(unique_time_array.begin(), unique_time_array.end()), unique_time_array.end());
        cout << "total # of unique time stamps: " << unique_time_array.size() << endl;
        const double learn_rate = curr_exp.get_online_learn_rate(), momentum = curr_exp.get_online_momentum();
        const vector<double> mode_parameter = curr_exp.get_test_mode_param(fold, kernel_clust_alg);
        vector<double> best_parameter(mode_parameter), delta_parameter(mode_parameter.size(), 0.0);
        pptr->init_test_prior(curr_exp.get_kernel_index(), test_kernel_param, mode_parameter);
        medgp_ctx *ctx = medgp_backend::context(test_kernel_param[0], test_kernel_param[1], test_kernel_param[2]);

        vector<HeldOut> tasks;          // in the reference's output order (tt major, jj minor)
        vector<double> mean, var;
        vector<int> status;
        vector<vector<double> > theta_of_task;  // only with updates
        float last_update_time = unique_time_array[0];
        for (int tt = 0; tt < (int)unique_time_array.size(); tt++) {
            const float stamp = unique_time_array[tt];
            vector<int> past_m, curr_m, curr_i;
            vector<float> past_t, past_v, curr_t, curr_v;
            for (size_t ii = 0; ii < time_array.size(); ii++) {
                if (time_array[ii] < stamp) {
                    if (!flag_update || fabs(time_array[ii] - stamp) <= 72.0) {  // 72 h history with updates
                        past_m.push_back(meta_array[ii]);
                        past_t.push_back(time_array[ii]);
                        past_v.push_back(value_array[ii]);
                    }
                } else if (time_array[ii] == stamp) {
                    curr_m.push_back(meta_array[ii]);
                    curr_i.push_back((int)ii);
                    curr_t.push_back(time_array[ii]);
                    curr_v.push_back(value_array[ii]);
                }
            }
            if (flag_update && (tt > 3) && (stamp - last_update_time) > 5.0 / 60.0) {
                // one momentum-SGD step on the past window (main_one_test.cpp:309-348)
                last_update_time = stamp;
                c_objective_one curr_objfunc(curr_exp.get_kernel_index(), test_kernel_param, past_m, past_t, past_v);
                double best_loss;
                vector<double> best_grads;
                const bool obj_flag = curr_objfunc.compute_objective(true, best_parameter, best_loss, best_grads,
                                                                     kptr, mptr, lptr, iptr, pptr);
                if (obj_flag) {
                    for (size_t h = 0; h < mode_parameter.size(); h++) {
                        const bool prior_flag = pptr->get_one_prior_flag((int)h);
                        const int prior_type = pptr->get_one_prior_type((int)h);
                        if ((!prior_flag) | (prior_type != 0)) {
                            delta_parameter[h] = momentum * delta_parameter[h] + learn_rate * best_grads[h];
                            best_parameter[h] -= delta_parameter[h];
                        }
                    }
                } else {
                    cout << "Warning: failed to update at t[" << tt << "] = " << stamp << "; reset to mode parameters" << endl;
                    best_parameter = mode_parameter;
                    std::fill(delta_parameter.begin(), delta_parameter.end(), 0.0);
                }
            }
            const size_t first_task = tasks.size();
            for (size_t jj = 0; jj < curr_t.size(); jj++) {
                HeldOut h;
                h.has_training = past_t.size() + curr_t.size() > 1;
                h.index = curr_i[jj];
                h.test_meta = curr_m[jj]; h.test_time = curr_t[jj]; h.test_value = curr_v[jj]; h.stamp = stamp;
                tasks.push_back(h);
            }
            mean.resize(tasks.size(), 0.0);
            var.resize(tasks.size(), 0.0);
            status.resize(tasks.size(), -1);
            if (flag_update)
                impute_time_stamp(ctx, best_parameter, past_m, past_t, past_v, curr_m, curr_t, curr_v, tasks, first_task,
                                  mean, var, status);
            if ((tt % 100) == 0) cout << "finish testing " << tt << "/" << unique_time_array.size() << " time stamps" << endl;
        }
        if (!flag_update) impute_without_update(ctx, best_parameter, meta_array, time_array, value_array, tasks, mean, var, status);
        write_imputation_outputs(curr_exp, output_prefix, PAN, tasks, mean, var, status, mode_parameter);
    }
    curr_exp.output_int_txt(curr_exp.get_exp_test_dir() + "test_" + output_prefix + "_flag_" + PAN,
                            vector<int>(1, (int)test_flag));
    cout << "finish (" << output_prefix << ") testing individual PAN " << PAN << " w/ " << time_array.size()
         << " samples; flag = " << test_flag << "; elapsed time = " << now_s() - t1 << " seconds" << endl;
}

int main(int argc, const char *argv[])
{
    if (argc != CMD_NUM) {
        cout << "ERROR: incorrect number of argument received!" << endl
             << "expect " << CMD_NUM << " but received " << argc << endl
             << "usage:\n\t --cfg\t --pan\t --thread\t --fold\t --kernclust-alg" << endl;
        return 1;
    }
    string exp_cfg, patient_PAN, kernel_clust_alg;
    int thread_num = 1, patient_fold = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--cfg")) exp_cfg = argv[++i];
        else if (!strcmp(argv[i], "--pan")) patient_PAN = argv[++i];
        else if (!strcmp(argv[i], "--thread")) thread_num = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fold")) patient_fold = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--kernclust-alg")) kernel_clust_alg = argv[++i];
        else { cout << "Error: unknown argument: " << argv[i] << endl; return 1; }
    }
    c_experiment curr_exp(exp_cfg);
    if (curr_exp.get_kernel_index() != 7) {
        cout << "Error: not supported kernel type " << curr_exp.get_kernel_index() << " (GPU backend: LMC-SM only)" << endl;
        return 1;
    }
    const double t1 = now_s();
    const vector<int> test_kernel_param = curr_exp.get_test_kernel_param(patient_fold, kernel_clust_alg);
    cout << "# of mixture for testing: " << test_kernel_param[0] << endl;
    c_kernel_LMC_SM kernel(test_kernel_param);
    c_inference_prior inffunc(thread_num);
    c_meanfunc_zero meanfunc;
    c_likelihood_gaussianMO likfunc(curr_exp.get_lik_param());
    c_prior prior(curr_exp.get_test_cov_num(patient_fold, kernel_clust_alg), curr_exp.get_mean_num(), curr_exp.get_lik_num());
    c_kernel *kptr = &kernel;
    c_meanfunc *mptr = &meanfunc;
    c_likelihood *lptr = &likfunc;
    c_inference *iptr = &inffunc;
    c_prior *pptr = &prior;
    run_test_one(curr_exp, patient_PAN, patient_fold, false, "mean_wo_update", kernel_clust_alg, kptr, mptr, lptr, iptr, pptr);
    run_test_one(curr_exp, patient_PAN, patient_fold, true, "mean_w_update", kernel_clust_alg, kptr, mptr, lptr, iptr, pptr);
    cout << "Finish all jobs. Total elapsed time = " << now_s() - t1 << " seconds" << endl;
    medgp_backend::shutdown();
    return 0;
}
