// imputation.h -- the pieces of the online-imputation test that main_one_test and
// main_cohort_test share: the held-out task list in the reference's output order, the
// per-observation refit path (medgp_cuda_predict, with jitter retries), the one-factorisation
// path (medgp_cuda_predict_online) with its fallback, and the output files
// (medgpc/src/main_one_test.cpp:269-472).
#ifndef MEDGP_IMPUTATION_H
#define MEDGP_IMPUTATION_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "c_experiment.h"
#include "medgp_host.h"

namespace {
using std::vector;

struct HeldOut {            // one imputation task
    vector<int> meta;       // training set (filled when the task is predicted on its own)
    vector<float> time, value;
    int test_meta, index;   // index: position of the held-out observation in the patient's arrays
    float test_time, test_value, stamp;
    bool has_training;
};

// training set of a task without hyper-parameter updates: every earlier observation, then the
// other observations sharing its time stamp (main_one_test.cpp:286-306, :354-366)
inline void fill_training(HeldOut &h, const vector<int> &meta_array, const vector<float> &time_array,
                   const vector<float> &value_array)
{
    h.meta.clear(); h.time.clear(); h.value.clear();
    for (int pass = 0; pass < 2; pass++)
        for (size_t ii = 0; ii < time_array.size(); ii++) {
            const bool take = pass == 0 ? time_array[ii] < h.stamp : (time_array[ii] == h.stamp && (int)ii != h.index);
            if (take) {
                h.meta.push_back(meta_array[ii]);
                h.time.push_back(time_array[ii]);
                h.value.push_back(value_array[ii]);
            }
        }
}

// predicts every task with `theta`; fills pred / ok.  Tasks without training data stay !ok.
inline void predict_tasks(medgp_ctx *ctx, const vector<double> &theta, const vector<HeldOut> &tasks, size_t b, size_t e,
                   vector<double> &mean, vector<double> &var, vector<int> &status)
{
    const size_t B = e - b;
    vector<int> sids, offs(1, 0), mstar;
    vector<float> xstar;
    vector<double> thetas;
    vector<size_t> which;
    for (size_t k = b; k < e; k++) {
        status[k] = -1;
        if (tasks[k].time.empty()) continue;
        int id = -1;
        if (medgp_cuda_add_series(ctx, (int)tasks[k].time.size(), (const int32_t *)tasks[k].meta.data(),
                                  tasks[k].time.data(), tasks[k].value.data(), &id) != MEDGP_OK) {
            std::cerr << "ERROR: medgp_cuda_add_series: " << medgp_cuda_last_error(ctx) << std::endl;
            exit(1);
        }
        sids.push_back(id);
        which.push_back(k);
        mstar.push_back(tasks[k].test_meta);
        xstar.push_back(tasks[k].test_time);
        offs.push_back((int)mstar.size());
        thetas.insert(thetas.end(), theta.begin(), theta.end());
    }
    (void)B;
    if (sids.empty()) return;
    vector<double> m(sids.size()), v(sids.size());
    vector<int> st(sids.size());
    if (medgp_cuda_predict(ctx, (int)sids.size(), sids.data(), thetas.data(), offs.data(), mstar.data(),
                           xstar.data(), m.data(), v.data(), st.data()) != MEDGP_OK) {
        std::cerr << "ERROR: medgp_cuda_predict: " << medgp_cuda_last_error(ctx) << std::endl;
        exit(1);
    }
    for (size_t q = 0; q < sids.size(); q++) {
        mean[which[q]] = (double)(float)m[q];  // the reference returns float moments
        var[which[q]] = (double)(float)v[q];
        status[which[q]] = st[q];
        medgp_cuda_free_series(ctx, sids[q]);
    }
}

// Tasks of a patient without updates, in the reference's order (time stamp major, array order
// minor); training sets are not materialised.
inline void list_tasks_without_update(const vector<int> &meta_array, const vector<float> &time_array,
                                      const vector<float> &value_array, vector<HeldOut> &tasks)
{
    vector<int> order(time_array.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return time_array[a] < time_array[b]; });
    for (size_t q = 0; q < order.size(); q++) {
        const int ii = order[q];
        HeldOut h;
        h.index = ii;
        h.test_meta = meta_array[ii]; h.test_time = time_array[ii]; h.test_value = value_array[ii];
        h.stamp = time_array[ii];
        // training data exist unless this is the only observation at the very first time stamp
        const bool first_stamp = time_array[ii] == time_array[order[0]];
        const bool shared = (q + 1 < order.size() && time_array[order[q + 1]] == time_array[ii]) ||
                            (q > 0 && time_array[order[q - 1]] == time_array[ii]);
        h.has_training = !first_stamp || shared;
        tasks.push_back(h);
    }
}

// mean / var / status of every task from the one-factorisation results of the patient
inline void scatter_online(const vector<HeldOut> &tasks, const double *m, const double *v, vector<double> &mean,
                           vector<double> &var, vector<int> &status)
{
    for (size_t k = 0; k < tasks.size(); k++) {
        if (!tasks[k].has_training) continue;  // status stays -1: the reference's zero-mean branch
        mean[k] = (double)(float)m[tasks[k].index];  // the reference returns float moments
        var[k] = (double)(float)v[tasks[k].index];
        status[k] = 0;
    }
}

inline void refit_per_observation(medgp_ctx *ctx, const vector<double> &theta, const vector<int> &meta_array,
                                  const vector<float> &time_array, const vector<float> &value_array,
                                  vector<HeldOut> &tasks, vector<double> &mean, vector<double> &var, vector<int> &status)
{
    const size_t step = 512;  // training sets per library call
    for (size_t b = 0; b < tasks.size(); b += step) {
        const size_t e = std::min(tasks.size(), b + step);
        for (size_t k = b; k < e; k++) fill_training(tasks[k], meta_array, time_array, value_array);
        predict_tasks(ctx, theta, tasks, b, e, mean, var, status);
        for (size_t k = b; k < e; k++) { tasks[k].meta.clear(); tasks[k].time.clear(); tasks[k].value.clear(); }
    }
}

// Without updates every training set is "the earlier observations plus the rest of the time
// stamp": ONE factorisation of the time-ordered patient serves all of them
// (medgp_cuda_predict_online).  If that matrix is not positive definite, or a time stamp holds
// too many observations, fall back to one training set per observation, which retries with
// jitter exactly as the reference does.
// MEDGP_NO_ONLINE=1 forces the reference's literal procedure (one fit per observation) everywhere;
// used to validate the one-factorisation paths against it.
inline bool online_paths_enabled()
{
    const char *ev = getenv("MEDGP_NO_ONLINE");
    return !(ev && atoi(ev) != 0);
}

inline void impute_without_update(medgp_ctx *ctx, const vector<double> &theta, const vector<int> &meta_array,
                                  const vector<float> &time_array, const vector<float> &value_array,
                                  vector<HeldOut> &tasks, vector<double> &mean, vector<double> &var, vector<int> &status)
{
    bool online = false;
    int sid = -1;
    if (online_paths_enabled() &&
        medgp_cuda_add_series_ordered(ctx, (int)time_array.size(), (const int32_t *)meta_array.data(), time_array.data(),
                                      value_array.data(), MEDGP_ORDER_TIME, &sid) == MEDGP_OK) {
        vector<double> m(time_array.size()), v(time_array.size());
        int st = -1;
        if (medgp_cuda_predict_online(ctx, 1, &sid, theta.data(), m.data(), v.data(), &st) != MEDGP_OK) {
            std::cerr << "ERROR: medgp_cuda_predict_online: " << medgp_cuda_last_error(ctx) << std::endl;
            exit(1);
        }
        medgp_cuda_free_series(ctx, sid);
        if (st == 0) {
            online = true;
            scatter_online(tasks, m.data(), v.data(), mean, var, status);
        }
    }
    if (!online) {
        std::cout << "Warning: one-factorisation imputation not applicable; refitting per observation" << std::endl;
        refit_per_observation(ctx, theta, meta_array, time_array, value_array, tasks, mean, var, status);
    }
}

// One time stamp of the with-update mode: tasks [first, end) are the observations `curr` of the
// stamp, each trained on `past` (the 72 h history) plus the other observations of the stamp
// (main_one_test.cpp:354-366).  With two or more observations one factorisation of past + curr
// serves them all; otherwise, or if it fails, every training set is fitted on its own.
inline void impute_time_stamp(medgp_ctx *ctx, const vector<double> &theta, const vector<int> &past_m,
                              const vector<float> &past_t, const vector<float> &past_v, const vector<int> &curr_m,
                              const vector<float> &curr_t, const vector<float> &curr_v, vector<HeldOut> &tasks,
                              size_t first, vector<double> &mean, vector<double> &var, vector<int> &status)
{
    const size_t g = curr_t.size(), np = past_t.size();
    if (g >= 2 && online_paths_enabled()) {
        vector<int> m(past_m);
        vector<float> t(past_t), v(past_v);
        m.insert(m.end(), curr_m.begin(), curr_m.end());
        t.insert(t.end(), curr_t.begin(), curr_t.end());
        v.insert(v.end(), curr_v.begin(), curr_v.end());
        int sid = -1;
        if (medgp_cuda_add_series_ordered(ctx, (int)t.size(), (const int32_t *)m.data(), t.data(), v.data(),
                                          MEDGP_ORDER_TIME, &sid) == MEDGP_OK) {
            vector<double> pm(t.size()), pv(t.size());
            int st = -1;
            if (medgp_cuda_predict_online(ctx, 1, &sid, theta.data(), pm.data(), pv.data(), &st) != MEDGP_OK) {
                std::cerr << "ERROR: medgp_cuda_predict_online: " << medgp_cuda_last_error(ctx) << std::endl;
                exit(1);
            }
            medgp_cuda_free_series(ctx, sid);
            if (st == 0) {
                for (size_t jj = 0; jj < g; jj++) {
                    mean[first + jj] = (double)(float)pm[np + jj];  // the reference returns float moments
                    var[first + jj] = (double)(float)pv[np + jj];
                    status[first + jj] = 0;
                }
                return;
            }
        }
    }
    for (size_t jj = 0; jj < g; jj++) {
        HeldOut &h = tasks[first + jj];
        h.meta = past_m; h.time = past_t; h.value = past_v;
        for (size_t kk = 0; kk < g; kk++)
            if (kk != jj) {  // same-time observations of other covariates join the training set
                h.meta.push_back(curr_m[kk]);
                h.time.push_back(curr_t[kk]);
                h.value.push_back(curr_v[kk]);
            }
    }
    predict_tasks(ctx, theta, tasks, first, first + g, mean, var, status);
    for (size_t jj = 0; jj < g; jj++) { tasks[first + jj].meta.clear(); tasks[first + jj].time.clear(); tasks[first + jj].value.clear(); }
}

// test_<mode>_{feature,ci}_<PAN>.txt and _{etime,error,pred}_<PAN>.bin (main_one_test.cpp:400-472)
inline void write_imputation_outputs(c_experiment &curr_exp, const std::string &output_prefix, const std::string &PAN,
                                     const vector<HeldOut> &tasks, const vector<double> &mean, const vector<double> &var,
                                     const vector<int> &status, const vector<double> &mode_parameter)
{
    vector<int> out_feature, out_ci;
    vector<double> out_etime, out_error, out_pred;
    for (size_t k = 0; k < tasks.size(); k++) {
        double impute_error;
        int ci;
        if (status[k] >= 0) {
            out_pred.push_back(mean[k]);
            impute_error = (float)mean[k] - tasks[k].test_value;
            ci = fabs(impute_error) <= 1.96 * sqrt(var[k]) ? 1 : 0;
        } else {
            // no training data or factorisation failure: zero-mean fallback (main_one_test.cpp:411-438)
            std::cout << "Warning: predict with zero mean for task " << k << std::endl;
            out_pred.push_back(0.0);
            impute_error = 0.0 - tasks[k].test_value;
            const double prior_var = exp(mode_parameter[tasks[k].test_meta]);
            ci = fabs(impute_error) <= 1.96 * prior_var ? 1 : 0;
        }
        out_error.push_back(impute_error);
        out_ci.push_back(ci);
        out_feature.push_back(curr_exp.get_feature_index()[tasks[k].test_meta]);
        out_etime.push_back(tasks[k].test_time - tasks[k].stamp);
    }
    if (!out_pred.empty()) {
        const std::string prefix = curr_exp.get_exp_test_dir() + "test_" + output_prefix + "_";
        curr_exp.output_int_txt(prefix + "feature_" + PAN, out_feature);
        curr_exp.output_double_bin(prefix + "etime_" + PAN, out_etime);
        curr_exp.output_int_txt(prefix + "ci_" + PAN, out_ci);
        curr_exp.output_double_bin(prefix + "error_" + PAN, out_error);
        curr_exp.output_double_bin(prefix + "pred_" + PAN, out_pred);
    }
}
}  // namespace
#endif  // MEDGP_IMPUTATION_H
