// main_cohort_test -- online imputation of a whole cohort, or one shard of it, on one GPU, without
// and with online hyper-parameter updates (main_one_test.cpp:140-141 runs both for one patient).
//   main_cohort_test --cfg exp_setup.json --pans <file with one PAN per line> --fold F
//                    --kernclust-alg A [--device d] [--shard i/N] [--update no|yes|both]
// Per patient this writes exactly the test_mean_wo_update_* / test_mean_w_update_* files
// main_one_test writes.  All patients of a fold share the mode hyper-parameters
// (kernel/fold<F>/<A>_mode_param.bin, c_experiment.cpp:198), and without updates one factorisation
// of the time-ordered patient yields every held-out prediction (medgp_cuda_predict_online), so the
// shard is ONE batched library call; patients for which that path does not apply are refitted
// per observation as main_one_test does.
// With updates theta moves along each patient's time axis (momentum SGD on the 72 h window,
// main_one_test.cpp:309-348), so time stays sequential PER PATIENT -- but patients are independent:
// the shard advances in lock-step over the patients' time-stamp index, and every super-step is two
// batched library calls for all patients at once: NLML+gradient on the update windows
// (medgp_cuda_nlml_grad), then one factorisation of "72 h history + this stamp" per patient whose
// leave-one-out identities give the predictions of the stamp's observations
// (medgp_cuda_predict_online).  The windows of a super-step are uploaded in one call
// (medgp_cuda_add_series_batch).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>

#include "c_experiment.h"
#include "imputation.h"
#include "medgp_host.h"

using std::cout;
using std::endl;
using std::string;
using std::vector;

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

namespace {
struct TestPatient {
    string pan;
    vector<int> meta;
    vector<float> time, value;
    int series_id = -1;
    vector<HeldOut> tasks;
    vector<double> mean, var;
    vector<int> status;
};
}  // namespace

namespace {
// One patient of the with-update mode: its position on its own time axis and its own theta.
struct UpdPatient {
    TestPatient *p = nullptr;
    vector<float> stamps;              // unique sorted time stamps
    vector<double> best, delta;        // current hyper-parameters, momentum buffer
    float last_update = 0.f;
    vector<HeldOut> tasks;             // in the reference's output order (tt major, jj minor)
    vector<double> mean, var;
    vector<int> status;
};

// With online updates (main_one_test.cpp:269-444 with flag_update): lock-step over the time-stamp
// index of all patients of the shard.
void run_with_update(medgp_ctx *ctx, c_experiment &curr_exp, const vector<TestPatient *> &mine, const vector<int> &kp,
                     const vector<double> &mode_parameter, int fold, const string &alg)
{
    const double t0 = now_s();
    const string output_prefix = "mean_w_update";
    const double learn_rate = curr_exp.get_online_learn_rate(), momentum = curr_exp.get_online_momentum();
    // the test-time prior table only clamps the A entries the mode kernel has at exactly 0
    // (c_prior.cpp:118-140); it depends on the mode parameters alone, so all patients share it
    c_prior prior(curr_exp.get_test_cov_num(fold, alg), curr_exp.get_mean_num(), curr_exp.get_lik_num());
    prior.init_test_prior(curr_exp.get_kernel_index(), kp, mode_parameter);
    c_objective_batch batch(ctx, kp[0], kp[1], kp[2]);
    // entries the SGD step moves: everything but the clamped ones (main_one_test.cpp:328-334)
    vector<char> movable(mode_parameter.size());
    for (size_t h = 0; h < mode_parameter.size(); h++) {
        const bool prior_flag = prior.get_one_prior_flag((int)h);
        const int prior_type = prior.get_one_prior_type((int)h);
        movable[h] = ((!prior_flag) | (prior_type != 0)) ? 1 : 0;
    }
    vector<UpdPatient> pats;
    size_t max_stamps = 0;
    for (TestPatient *p : mine) {
        if (p->time.empty()) continue;
        UpdPatient u;
        u.p = p;
        u.stamps = p->time;
        std::sort(u.stamps.begin(), u.stamps.end());
        u.stamps.erase(std::unique(u.stamps.begin(), u.stamps.end()), u.stamps.end());
        u.best = mode_parameter;
        u.delta.assign(mode_parameter.size(), 0.0);
        u.last_update = u.stamps[0];
        max_stamps = std::max(max_stamps, u.stamps.size());
        pats.push_back(std::move(u));
    }
    long n_updates = 0, n_failed_updates = 0, n_predictions = 0, n_refits = 0;
    double t_win = 0, t_up_a = 0, t_eval = 0, t_sgd = 0, t_up_b = 0, t_pred = 0;  // where a super-step goes (printed at the end)
    struct Window { vector<int> past_m, curr_m, curr_i; vector<float> past_t, past_v, curr_t, curr_v; };
    vector<Window> win(pats.size());
    for (size_t tt = 0; tt < max_stamps; tt++) {
        // ---- the windows of this super-step (main_one_test.cpp:286-306)
        double tp = now_s();
        vector<size_t> act;
        for (size_t k = 0; k < pats.size(); k++)
            if (tt < pats[k].stamps.size()) act.push_back(k);
#pragma omp parallel for schedule(dynamic, 16)
        for (long a = 0; a < (long)act.size(); a++) {
            const size_t k = act[a];
            UpdPatient &u = pats[k];
            Window &w = win[k];
            w = Window();
            const float stamp = u.stamps[tt];
            const TestPatient &p = *u.p;
            for (size_t ii = 0; ii < p.time.size(); ii++) {
                if (p.time[ii] < stamp) {
                    if (fabs(p.time[ii] - stamp) <= 72.0) {  // 72 h history
                        w.past_m.push_back(p.meta[ii]); w.past_t.push_back(p.time[ii]); w.past_v.push_back(p.value[ii]);
                    }
                } else if (p.time[ii] == stamp) {
                    w.curr_m.push_back(p.meta[ii]); w.curr_i.push_back((int)ii);
                    w.curr_t.push_back(p.time[ii]); w.curr_v.push_back(p.value[ii]);
                }
            }
        }
        if (act.empty()) break;
        t_win += now_s() - tp;
        // ---- one upload per super-step: the 72 h history of every patient that needs it, for the
        //      update (more than 2 points, c_objective_one.cpp:51) and/or as the training set of
        //      a stamp with a single observation
        vector<size_t> upd;              // patients with an update due at this stamp
        vector<long> past_sid(pats.size(), -1);
        {
            tp = now_s();
            vector<size_t> need;
            vector<int> ns;
            vector<int32_t> cm;
            vector<float> cx, cy;
            for (size_t k : act) {
                UpdPatient &u = pats[k];
                const Window &w = win[k];
                const float stamp = u.stamps[tt];
                const bool due = (tt > 3) && (stamp - u.last_update) > 5.0 / 60.0;
                if (due) {
                    u.last_update = stamp;
                    upd.push_back(k);
                }
                const bool for_update = due && w.past_t.size() > 2;
                const bool for_predict = w.curr_t.size() == 1 && !w.past_t.empty() && online_paths_enabled();
                if (!for_update && !for_predict) continue;
                need.push_back(k);
                ns.push_back((int)w.past_t.size());
                cm.insert(cm.end(), w.past_m.begin(), w.past_m.end());
                cx.insert(cx.end(), w.past_t.begin(), w.past_t.end());
                cy.insert(cy.end(), w.past_v.begin(), w.past_v.end());
            }
            if (!need.empty()) {
                vector<int> sids(need.size());
                if (medgp_cuda_add_series_batch(ctx, (int)need.size(), ns.data(), cm.data(), cx.data(), cy.data(),
                                                MEDGP_ORDER_FEATURE, sids.data()) != MEDGP_OK) {
                    std::cerr << "ERROR: medgp_cuda_add_series_batch: " << medgp_cuda_last_error(ctx) << endl;
                    exit(1);
                }
                for (size_t q = 0; q < need.size(); q++) past_sid[need[q]] = sids[q];
            }
            t_up_a += now_s() - tp;
        }
        // ---- (a) one momentum-SGD step on the past window where one is due (main_one_test.cpp:309-348)
        {
            tp = now_s();
            vector<size_t> evaluated;  // ... whose window can be evaluated
            vector<medgp_eval_request> reqs;
            for (size_t k : upd)
                if (win[k].past_t.size() > 2) {
                    evaluated.push_back(k);
                    reqs.push_back({(int)past_sid[k], &pats[k].best, &prior});
                }
            vector<medgp_eval_result> res;
            if (!reqs.empty()) batch.compute(true, reqs, res);
            t_eval += now_s() - tp;
            tp = now_s();
            vector<long> slot(pats.size(), -1);  // patient -> its result, or -1: window too small
            for (size_t q = 0; q < evaluated.size(); q++) slot[evaluated[q]] = (long)q;
#pragma omp parallel for schedule(static)
            for (long a = 0; a < (long)upd.size(); a++) {
                UpdPatient &u = pats[upd[a]];
                const long q = slot[upd[a]];
                if (q >= 0 && res[q].ok) {
                    const vector<double> &g = res[q].grad;
                    for (size_t h = 0; h < mode_parameter.size(); h++)
                        if (movable[h]) {
                            u.delta[h] = momentum * u.delta[h] + learn_rate * g[h];
                            u.best[h] -= u.delta[h];
                        }
                } else {
                    u.best = mode_parameter;
                    std::fill(u.delta.begin(), u.delta.end(), 0.0);
                }
            }
            for (size_t k : upd) {
                const long q = slot[k];
                if (q >= 0 && res[q].ok) {
                    n_updates++;
                } else {
                    cout << "Warning: PAN " << pats[k].p->pan << ": failed to update at t[" << tt << "] = " << pats[k].stamps[tt]
                         << "; reset to mode parameters" << endl;
                    n_failed_updates++;
                }
            }
            t_sgd += now_s() - tp;
        }
        // ---- (b) the observations of the stamp, each from "72 h history + the rest of the stamp"
        //      (main_one_test.cpp:352-399)
        vector<size_t> first_task(pats.size(), 0);
        vector<size_t> fit;        // patients with training data at this stamp
        vector<size_t> single;     // ... a single observation: predicted from the uploaded history
        vector<size_t> multi;      // ... several: one factorisation of history + stamp (leave-one-out inside the stamp)
        vector<int> ns;
        vector<int32_t> cm;
        vector<float> cx, cy;
        for (size_t k : act) {
            UpdPatient &u = pats[k];
            const Window &w = win[k];
            first_task[k] = u.tasks.size();
            for (size_t jj = 0; jj < w.curr_t.size(); jj++) {
                HeldOut h;
                h.has_training = w.past_t.size() + w.curr_t.size() > 1;
                h.index = w.curr_i[jj];
                h.test_meta = w.curr_m[jj]; h.test_time = w.curr_t[jj]; h.test_value = w.curr_v[jj]; h.stamp = u.stamps[tt];
                u.tasks.push_back(h);
            }
            u.mean.resize(u.tasks.size(), 0.0);
            u.var.resize(u.tasks.size(), 0.0);
            u.status.resize(u.tasks.size(), -1);
            if (w.past_t.size() + w.curr_t.size() <= 1) continue;  // no training data: zero-mean fallback at output
            fit.push_back(k);
            if (w.curr_t.size() == 1) {
                if (past_sid[k] >= 0) single.push_back(k);
                continue;
            }
            multi.push_back(k);
            ns.push_back((int)(w.past_t.size() + w.curr_t.size()));
            cm.insert(cm.end(), w.past_m.begin(), w.past_m.end()); cm.insert(cm.end(), w.curr_m.begin(), w.curr_m.end());
            cx.insert(cx.end(), w.past_t.begin(), w.past_t.end()); cx.insert(cx.end(), w.curr_t.begin(), w.curr_t.end());
            cy.insert(cy.end(), w.past_v.begin(), w.past_v.end()); cy.insert(cy.end(), w.curr_v.begin(), w.curr_v.end());
        }
        vector<char> done(pats.size(), 0);
        const size_t P = mode_parameter.size();
        if (!single.empty()) {  // GP_Regression::train(false) + predict of one point (main_one_test.cpp:386-399), batched
            tp = now_s();
            vector<int> sids(single.size()), offs(single.size() + 1, 0), st(single.size(), -1);
            vector<int32_t> mstar(single.size());
            vector<float> xstar(single.size());
            vector<double> thetas(single.size() * P), m(single.size()), v(single.size());
#pragma omp parallel for schedule(static)
            for (long q = 0; q < (long)single.size(); q++) {
                const size_t k = single[q];
                sids[q] = (int)past_sid[k];
                offs[q + 1] = (int)q + 1;
                mstar[q] = win[k].curr_m[0];
                xstar[q] = win[k].curr_t[0];
                std::copy(pats[k].best.begin(), pats[k].best.end(), thetas.begin() + (size_t)q * P);
            }
            if (medgp_cuda_predict(ctx, (int)single.size(), sids.data(), thetas.data(), offs.data(), mstar.data(), xstar.data(),
                                   m.data(), v.data(), st.data()) != MEDGP_OK) {
                std::cerr << "ERROR: medgp_cuda_predict: " << medgp_cuda_last_error(ctx) << endl;
                exit(1);
            }
            for (size_t q = 0; q < single.size(); q++) {
                if (st[q] < 0) continue;
                UpdPatient &u = pats[single[q]];
                u.mean[first_task[single[q]]] = (double)(float)m[q];  // the reference returns float moments
                u.var[first_task[single[q]]] = (double)(float)v[q];
                u.status[first_task[single[q]]] = st[q];
                done[single[q]] = 1;
                n_predictions++;
            }
            t_pred += now_s() - tp;
        }
        if (!multi.empty() && online_paths_enabled()) {
            vector<int> sids(multi.size());
            tp = now_s();
            const int rc_up = medgp_cuda_add_series_batch(ctx, (int)multi.size(), ns.data(), cm.data(), cx.data(), cy.data(),
                                                          MEDGP_ORDER_TIME, sids.data());
            t_up_b += now_s() - tp;
            if (rc_up == MEDGP_OK) {
                tp = now_s();
                size_t ntot = 0;
                vector<double> thetas(multi.size() * P);
                for (size_t q = 0; q < multi.size(); q++) ntot += (size_t)ns[q];
                for (size_t q = 0; q < multi.size(); q++)
                    std::copy(pats[multi[q]].best.begin(), pats[multi[q]].best.end(), thetas.begin() + q * P);
                vector<double> m(ntot), v(ntot);
                vector<int> st(multi.size(), -1);
                if (medgp_cuda_predict_online(ctx, (int)multi.size(), sids.data(), thetas.data(), m.data(), v.data(), st.data()) != MEDGP_OK) {
                    std::cerr << "ERROR: medgp_cuda_predict_online: " << medgp_cuda_last_error(ctx) << endl;
                    exit(1);
                }
                size_t off = 0;
                for (size_t q = 0; q < multi.size(); q++) {
                    UpdPatient &u = pats[multi[q]];
                    const Window &w = win[multi[q]];
                    const size_t np = w.past_t.size(), g = w.curr_t.size();
                    if (st[q] == 0) {
                        for (size_t jj = 0; jj < g; jj++) {
                            u.mean[first_task[multi[q]] + jj] = (double)(float)m[off + np + jj];  // the reference returns float moments
                            u.var[first_task[multi[q]] + jj] = (double)(float)v[off + np + jj];
                            u.status[first_task[multi[q]] + jj] = 0;
                        }
                        done[multi[q]] = 1;
                        n_predictions += (long)g;
                    }
                    off += (size_t)ns[q];
                }
                t_pred += now_s() - tp;
                tp = now_s();
                medgp_cuda_free_series_batch(ctx, (int)sids.size(), sids.data());
                t_up_b += now_s() - tp;
            }
        }
        {   // the histories of this super-step are done with
            tp = now_s();
            vector<int> sids;
            for (size_t k : act)
                if (past_sid[k] >= 0) sids.push_back((int)past_sid[k]);
            if (!sids.empty()) medgp_cuda_free_series_batch(ctx, (int)sids.size(), sids.data());
            t_up_a += now_s() - tp;
        }
        for (size_t k : fit) {  // whatever the batched path could not serve: the single-patient routine (refits, jitter)
            if (done[k]) continue;
            UpdPatient &u = pats[k];
            const Window &w = win[k];
            impute_time_stamp(ctx, u.best, w.past_m, w.past_t, w.past_v, w.curr_m, w.curr_t, w.curr_v, u.tasks, first_task[k],
                              u.mean, u.var, u.status);
            n_refits += (long)w.curr_t.size();
        }
    }
    batch.release();
    for (UpdPatient &u : pats)
        write_imputation_outputs(curr_exp, output_prefix, u.p->pan, u.tasks, u.mean, u.var, u.status, mode_parameter);
    for (TestPatient *p : mine)
        curr_exp.output_int_txt(curr_exp.get_exp_test_dir() + "test_" + output_prefix + "_flag_" + p->pan,
                                vector<int>(1, (int)!p->time.empty()));
    cout << "with updates: " << n_predictions << " predictions and " << n_updates << " hyper-parameter updates ("
         << n_failed_updates << " reset) in " << max_stamps << " lock-step super-steps, " << n_refits
         << " predictions by the single-patient routine; elapsed time = " << now_s() - t0 << " seconds" << endl;
    cout << "  of which: windows " << t_win << " s, window uploads " << t_up_a + t_up_b << " s, NLML+gradient calls " << t_eval
         << " s, SGD steps " << t_sgd << " s, imputation calls " << t_pred << " s" << endl;
}
}  // namespace

int main(int argc, const char *argv[])
{
    string exp_cfg, pan_file, kernel_clust_alg;
    int device = 0, shard = 0, nshard = 1, fold = 0;
    string update_arg = "both";
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--cfg") && i + 1 < argc) exp_cfg = argv[++i];
        else if (!strcmp(argv[i], "--pans") && i + 1 < argc) pan_file = argv[++i];
        else if (!strcmp(argv[i], "--fold") && i + 1 < argc) fold = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--kernclust-alg") && i + 1 < argc) kernel_clust_alg = argv[++i];
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--update") && i + 1 < argc) update_arg = argv[++i];
        else if (!strcmp(argv[i], "--shard") && i + 1 < argc) {
            if (sscanf(argv[++i], "%d/%d", &shard, &nshard) != 2 || nshard < 1 || shard < 0 || shard >= nshard) {
                cout << "Error: --shard expects i/N" << endl;
                return 1;
            }
        } else {
            cout << "usage: main_cohort_test --cfg exp_setup.json --pans list.txt --fold F --kernclust-alg A [--device d] [--shard i/N] [--update no|yes|both]" << endl;
            return 1;
        }
    }
    if (update_arg != "no" && update_arg != "yes" && update_arg != "both") {
        cout << "Error: --update expects no, yes or both" << endl;
        return 1;
    }
    if (exp_cfg.empty() || pan_file.empty() || kernel_clust_alg.empty()) {
        cout << "Error: --cfg, --pans and --kernclust-alg are required" << endl;
        return 1;
    }
    c_experiment curr_exp(exp_cfg);
    if (curr_exp.get_kernel_index() != 7) {
        cout << "Error: not supported kernel type " << curr_exp.get_kernel_index() << " (GPU backend: LMC-SM only)" << endl;
        return 1;
    }
    const vector<int> kp = curr_exp.get_test_kernel_param(fold, kernel_clust_alg);
    const vector<double> mode_parameter = curr_exp.get_test_mode_param(fold, kernel_clust_alg);
    const string output_prefix = "mean_wo_update";

    // ---- load the cohort, deal shards by descending n^3
    vector<string> pans;
    {
        std::ifstream f(pan_file.c_str());
        string line;
        while (f >> line) pans.push_back(line);
    }
    // sizes-only pass over the whole cohort, LPT deal, then only this shard's patients are parsed
    const vector<int> sizes = curr_exp.get_cohort_sizes(pans);
    const vector<int> shard_of = medgp_lpt_assign(sizes, nshard);
    vector<string> my_pans;
    for (size_t k = 0; k < pans.size(); k++)
        if (shard_of[k] == shard) my_pans.push_back(pans[k]);
    vector<c_experiment::patient_data> loaded;
    curr_exp.get_cohort_data(my_pans, loaded);
    vector<TestPatient> all(my_pans.size());
    vector<TestPatient *> mine;
    for (size_t k = 0; k < my_pans.size(); k++) {
        all[k].pan = my_pans[k];
        all[k].meta.swap(loaded[k].meta);
        all[k].time.swap(loaded[k].time);
        all[k].value.swap(loaded[k].value);
        mine.push_back(&all[k]);
    }
    cout << "shard " << shard << "/" << nshard << ": " << mine.size() << " of " << pans.size() << " patients on device " << device << endl;

    medgp_ctx *ctx = medgp_backend::context(kp[0], kp[1], kp[2], device);
    const double t0 = now_s();
    if (update_arg != "yes") {
    // ---- upload time-ordered, one batched call for the shard
    vector<int> sids;
    vector<TestPatient *> owner;
    size_t ntot = 0;
    for (TestPatient *p : mine) {
        if (p->time.empty()) continue;
        list_tasks_without_update(p->meta, p->time, p->value, p->tasks);
        p->mean.assign(p->tasks.size(), 0.0);
        p->var.assign(p->tasks.size(), 0.0);
        p->status.assign(p->tasks.size(), -1);
        if (!online_paths_enabled() ||
            medgp_cuda_add_series_ordered(ctx, (int)p->time.size(), (const int32_t *)p->meta.data(), p->time.data(),
                                          p->value.data(), MEDGP_ORDER_TIME, &p->series_id) != MEDGP_OK) {
            p->series_id = -1;  // e.g. too many observations on one time stamp: refit below
            continue;
        }
        sids.push_back(p->series_id);
        owner.push_back(p);
        ntot += p->time.size();
    }
    long predictions = 0, refits = 0;
    if (!sids.empty()) {
        vector<double> thetas, m(ntot), v(ntot);
        for (size_t k = 0; k < sids.size(); k++) thetas.insert(thetas.end(), mode_parameter.begin(), mode_parameter.end());
        vector<int> st(sids.size(), -1);
        if (medgp_cuda_predict_online(ctx, (int)sids.size(), sids.data(), thetas.data(), m.data(), v.data(), st.data()) != MEDGP_OK) {
            std::cerr << "ERROR: medgp_cuda_predict_online: " << medgp_cuda_last_error(ctx) << endl;
            return 1;
        }
        size_t off = 0;
        for (size_t k = 0; k < sids.size(); k++) {
            TestPatient *p = owner[k];
            if (st[k] == 0) {
                scatter_online(p->tasks, m.data() + off, v.data() + off, p->mean, p->var, p->status);
                predictions += (long)p->tasks.size();
            } else {
                p->series_id = -1;
            }
            off += p->time.size();
            medgp_cuda_free_series(ctx, sids[k]);
        }
    }
    const double t_online = now_s() - t0;
    for (TestPatient *p : mine) {
        if (p->time.empty() || p->series_id >= 0) continue;
        cout << "Warning: PAN " << p->pan << ": one-factorisation imputation not applicable; refitting per observation" << endl;
        refit_per_observation(ctx, mode_parameter, p->meta, p->time, p->value, p->tasks, p->mean, p->var, p->status);
        refits += (long)p->tasks.size();
    }
    // ---- outputs
    for (TestPatient *p : mine) {
        const bool test_flag = !p->time.empty();
        if (test_flag) write_imputation_outputs(curr_exp, output_prefix, p->pan, p->tasks, p->mean, p->var, p->status, mode_parameter);
        curr_exp.output_int_txt(curr_exp.get_exp_test_dir() + "test_" + output_prefix + "_flag_" + p->pan,
                                vector<int>(1, (int)test_flag));
    }
    cout << "without updates: " << predictions << " predictions from one factorisation per patient in " << t_online
         << " s, " << refits << " by per-observation refits; elapsed time = " << now_s() - t0 << " seconds" << endl;
    }
    if (update_arg != "no") run_with_update(ctx, curr_exp, mine, kp, mode_parameter, fold, kernel_clust_alg);
    cout << "Finish all jobs. Total elapsed time = " << now_s() - t0 << " seconds" << endl;
    medgp_backend::shutdown();
    return 0;
}
