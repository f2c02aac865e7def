// main_cohort_test -- online imputation (mode mean_wo_update) of a whole cohort, or one shard
// of it, on one GPU.
//   main_cohort_test --cfg exp_setup.json --pans <file with one PAN per line> --fold F
//                    --kernclust-alg A [--device d] [--shard i/N]
// Per patient this writes exactly the test_mean_wo_update_* files main_one_test writes (SURVEY.md
// appendix B).  All patients of a fold share the mode hyper-parameters
// (kernel/fold<F>/<A>_mode_param.bin, c_experiment.cpp:198), and without updates one factorisation
// of the time-ordered patient yields every held-out prediction (medgp_cuda_predict_online), so the
// shard is ONE batched library call; patients for which that path does not apply are refitted
// per observation as main_one_test does.  The with-update mode is sequential in time per patient
// and stays with main_one_test.
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

int main(int argc, const char *argv[])
{
    string exp_cfg, pan_file, kernel_clust_alg;
    int device = 0, shard = 0, nshard = 1, fold = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--cfg") && i + 1 < argc) exp_cfg = argv[++i];
        else if (!strcmp(argv[i], "--pans") && i + 1 < argc) pan_file = argv[++i];
        else if (!strcmp(argv[i], "--fold") && i + 1 < argc) fold = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--kernclust-alg") && i + 1 < argc) kernel_clust_alg = argv[++i];
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--shard") && i + 1 < argc) {
            if (sscanf(argv[++i], "%d/%d", &shard, &nshard) != 2 || nshard < 1 || shard < 0 || shard >= nshard) {
                cout << "Error: --shard expects i/N" << endl;
                return 1;
            }
        } else {
            cout << "usage: main_cohort_test --cfg exp_setup.json --pans list.txt --fold F --kernclust-alg A [--device d] [--shard i/N]" << endl;
            return 1;
        }
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
    vector<TestPatient> all(pans.size());
    for (size_t k = 0; k < pans.size(); k++) {
        all[k].pan = pans[k];
        curr_exp.get_one_patient_data(pans[k], all[k].meta, all[k].time, all[k].value);
    }
    vector<size_t> order(all.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return all[a].time.size() > all[b].time.size(); });
    vector<double> load(nshard, 0.0);
    vector<TestPatient *> mine;
    for (size_t k : order) {
        const int tgt = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        const double n = (double)all[k].time.size();
        load[tgt] += n * n * n;
        if (tgt == shard) mine.push_back(&all[k]);
    }
    cout << "shard " << shard << "/" << nshard << ": " << mine.size() << " of " << all.size() << " patients on device " << device << endl;

    medgp_ctx *ctx = medgp_backend::context(kp[0], kp[1], kp[2], device);
    const double t0 = now_s();
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
    cout << "Finish all jobs. " << predictions << " predictions from one factorisation per patient in " << t_online
         << " s, " << refits << " by per-observation refits; total elapsed time = " << now_s() - t0 << " seconds" << endl;
    medgp_backend::shutdown();
    return 0;
}
