// main_one_train -- train one patient (same CLI, flow and output files as the reference's
// medgpc/src/main_one_train.cpp:41-324), numerical work on the GPU through libmedgp_cuda.so.
//   main_one_train --cfg exp_setup.json --pan <PAN> --thread <T>
// Outputs in <exp_train_dir>/: train_init_hyp_<PAN>.bin, train_hyp_<PAN>.bin,
// train_var_hyp_<PAN>.bin (prior 2), train_num_<PAN>.txt, train_flag_<PAN>.txt.
#include <chrono>
#include <cstring>
#include <iostream>
#include <limits>

#include "c_experiment.h"
#include "medgp_host.h"

using std::cout;
using std::endl;
using std::string;
using std::vector;

#define CMD_NUM 7

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void run_train_one(c_experiment &curr_exp, const string &PAN, c_kernel *&kptr, c_meanfunc *&mptr,
                          c_likelihood *&lptr, c_inference *&iptr, c_prior *&pptr)
{
    iptr->print_inffunc();
    kptr->print_kernel();
    cout << "running individual training..." << endl << "current patinet PAN = " << PAN << endl;

    vector<vector<double> > global_hyp_array;
    curr_exp.get_global_hyp(global_hyp_array);

    vector<int> meta_array;
    vector<float> time_array, value_array;
    curr_exp.get_one_patient_data(PAN, meta_array, time_array, value_array);
    cout << "current number of data points = " << time_array.size() << endl;

    // every output needs at least two observations (main_one_train.cpp:181-197)
    const double t1 = now_s();
    vector<int> count_array(curr_exp.get_feature_index().size(), 0);
    for (size_t t = 0; t < time_array.size(); t++) count_array[meta_array[t]] += 1;
    bool sample_flag = true;
    for (size_t f = 0; f < count_array.size(); f++)
        if (count_array[f] < 2) { sample_flag = false; break; }

    bool flag_data = false;
    if (!sample_flag) {
        cout << "skip due to insufficient # of samples" << endl;
    } else {
        c_objective_one curr_objfunc(curr_exp.get_kernel_index(), curr_exp.get_kernel_param(), meta_array,
                                     time_array, value_array);
        c_objective *obj_ptr = &curr_objfunc;
        double best_loss = std::numeric_limits<double>::max();
        vector<double> best_init, opt_parameter;
        bool success = false;

        pptr->initialize_param(curr_exp.get_cov_num(), curr_exp.get_mean_num(), curr_exp.get_lik_num());
        cout << "finish initialization of prior" << endl;

        // random-initialisation scoring: NLML only, priors inactive, same data, different theta.
        // The reference loops (main_one_train.cpp:228-253); here all candidates go to the GPU as
        // ONE batch and the reference's selection rule (stop at the first failure) is replayed.
        const double t3 = now_s();
        const int n_init = curr_exp.get_scg_init_num();
        if ((int)time_array.size() > 2 && n_init > 0) {
            const vector<int> kp = curr_exp.get_kernel_param();
            medgp_ctx *ctx = medgp_backend::context(kp[0], kp[1], kp[2]);
            const int sid = medgp_backend::series_for(ctx, meta_array, time_array, value_array);
            c_objective_batch batch(ctx, kp[0], kp[1], kp[2]);
            vector<medgp_eval_request> reqs(n_init);
            for (int k = 0; k < n_init; k++) reqs[k] = {sid, &global_hyp_array[k], pptr};
            vector<medgp_eval_result> res;
            batch.compute(false, reqs, res);
            for (int init = 0; init < n_init; init++) {
                success = res[init].ok;
                if (!success) {
                    cout << "WARNING: failed in computing objective!" << endl;
                    break;
                }
                if (res[init].value < best_loss) {
                    best_loss = res[init].value;
                    best_init = global_hyp_array[init];
                }
            }
        }
        cout << "INFO: finish initialization " << n_init << "; time usage = " << now_s() - t3 << " seconds" << endl;
        curr_exp.output_double_bin(curr_exp.get_exp_train_dir() + "train_init_hyp_" + PAN, best_init);

        if (success) {
            pptr->setup_param(curr_exp.get_kernel_index(), curr_exp.get_kernel_param(),
                              curr_exp.get_prior_mode(), curr_exp.get_prior_hyp());
            const double t5 = now_s();
            cout << "start doing optimization" << endl;
            if (curr_exp.get_prior_mode() == 2) {  // sparse hierarchical-gamma prior: variational EM
                c_optimizer_varEM curr_optfunc;
                curr_optfunc.set_sub_opt_iter(curr_exp.get_prior_sub_opt_iter());
                curr_optfunc.optimize((-1) * curr_exp.get_scg_max_iter_num(), best_init, obj_ptr, true, best_loss,
                                      opt_parameter, kptr, mptr, lptr, iptr, pptr);
            } else {
                c_optimizer_scg curr_optfunc;
                curr_optfunc.optimize((-1) * curr_exp.get_scg_max_iter_num(), best_init, obj_ptr, false, best_loss,
                                      opt_parameter, kptr, mptr, lptr, iptr, pptr);
            }
            cout << "total time for doing optimization for patient " << PAN << " = " << now_s() - t5
                 << " seconds; final loss = " << best_loss << endl;
            curr_exp.output_double_bin(curr_exp.get_exp_train_dir() + "train_hyp_" + PAN, opt_parameter);
            if (curr_exp.get_prior_mode() == 2)
                curr_exp.output_double_bin(curr_exp.get_exp_train_dir() + "train_var_hyp_" + PAN,
                                           pptr->get_cov_varEM_all());
        }
        flag_data = success;
        cout << "finish individual id: " << PAN << " w/ " << time_array.size() << " samples; flag = " << flag_data
             << "; elapsed time = " << now_s() - t1 << " seconds" << endl;
    }
    curr_exp.output_int_txt(curr_exp.get_exp_train_dir() + "train_num_" + PAN, vector<int>(1, (int)time_array.size()));
    curr_exp.output_int_txt(curr_exp.get_exp_train_dir() + "train_flag_" + PAN, vector<int>(1, (int)flag_data));
}

int main(int argc, const char *argv[])
{
    if (argc != CMD_NUM) {
        cout << "ERROR: incorrect number of argument received!" << endl
             << "expect " << CMD_NUM << " but received " << argc << endl
             << "usage:" << endl
             << "\t --cfg:\t the JSON configuration file" << endl
             << "\t --pan:\t ID of the training patient" << endl
             << "\t --thread:\t kept for compatibility (the GPU backend ignores it)" << endl;
        return 1;
    }
    string exp_cfg, patient_PAN;
    int thread_num = 1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--cfg")) exp_cfg = argv[++i];
        else if (!strcmp(argv[i], "--pan")) patient_PAN = argv[++i];
        else if (!strcmp(argv[i], "--thread")) thread_num = atoi(argv[++i]);
        else { cout << "Error: unknown argument: " << argv[i] << endl; return 1; }
    }
    cout << "current configuration file: " << exp_cfg << endl
         << "current training patient: " << patient_PAN << endl
         << "current threading number for matrix operation: " << thread_num << endl;
    c_experiment curr_exp(exp_cfg);
    if (curr_exp.get_kernel_index() != 7) {
        cout << "Error: not supported kernel type " << curr_exp.get_kernel_index() << " (GPU backend: LMC-SM only)" << endl;
        return 1;
    }
    const double t1 = now_s();
    c_kernel_LMC_SM kernel(curr_exp.get_kernel_param());
    c_inference_prior inffunc(thread_num);
    c_meanfunc_zero meanfunc;
    c_likelihood_gaussianMO likfunc(curr_exp.get_lik_param());
    c_prior prior(curr_exp.get_cov_num(), curr_exp.get_mean_num(), curr_exp.get_lik_num());
    c_kernel *kptr = &kernel;
    c_meanfunc *mptr = &meanfunc;
    c_likelihood *lptr = &likfunc;
    c_inference *iptr = &inffunc;
    c_prior *pptr = &prior;
    run_train_one(curr_exp, patient_PAN, kptr, mptr, lptr, iptr, pptr);
    cout << "Finish all jobs. Total elapsed time = " << now_s() - t1 << " seconds" << endl;
    medgp_backend::shutdown();
    return 0;
}
