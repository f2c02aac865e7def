// main_cohort_train -- train a whole cohort (or one shard of it) on one GPU in lock-step.
//   main_cohort_train --cfg exp_setup.json --pans <file with one PAN per line>
//                     [--device d] [--shard i/N] [--resume] [--host-optimizer] [--max-evals E]
// --resume skips the patients whose train_flag_<PAN>.txt already says 1 (the reference's way of
// resuming is to re-submit the jobs of the patients without results).
// Per patient this produces exactly the files main_one_train writes and
// follows the same procedure (score random_init_num random initialisations, optimise the best
// with SCG or variational EM), but every optimiser super-step evaluates ALL active patients at
// once, and the line searches themselves run ON THE DEVICE (medgp_cuda_scg_*: one state machine
// per patient in HBM, theta never crosses PCIe; the host keeps the variational-EM rounds).
// --host-optimizer selects the older path: host steppers, one batched library call (theta up,
// gradients down) per super-step.  The reference deploys one job per patient
// (medgpc/util/run_exp_generator.py:213-260); sharding here is the same idea per GPU:
// patients are dealt to shards by descending n^3 (LPT), and shards never communicate.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>

#include "c_experiment.h"
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
struct Patient {
    string pan;
    vector<int> meta;
    vector<float> time, value;
    bool enough = false, success = false;
    int series_id = -1;
    vector<double> best_init, opt_parameter;
    double best_loss = std::numeric_limits<double>::max();
    std::unique_ptr<c_prior> prior;
    scg_stepper scg;
    varem_stepper vem;
    bool use_vem = false, active = false;
    bool wants_eval() const { return use_vem ? vem.wants_eval() : scg.wants_eval(); }
    const vector<double> &point() const { return use_vem ? vem.point() : scg.point(); }
    void feed(bool ok, double f, const vector<double> &g) { use_vem ? vem.feed(ok, f, g) : scg.feed(ok, f, g); }
};
}  // namespace

int main(int argc, const char *argv[])
{
    string exp_cfg, pan_file;
    int device = 0, shard = 0, nshard = 1;
    long max_evals = -1;
    bool host_optimizer = false, resume = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--cfg") && i + 1 < argc) exp_cfg = argv[++i];
        else if (!strcmp(argv[i], "--pans") && i + 1 < argc) pan_file = argv[++i];
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--max-evals") && i + 1 < argc) max_evals = atol(argv[++i]);
        else if (!strcmp(argv[i], "--host-optimizer")) host_optimizer = true;
        else if (!strcmp(argv[i], "--resume")) resume = true;
        else if (!strcmp(argv[i], "--shard") && i + 1 < argc) {
            if (sscanf(argv[++i], "%d/%d", &shard, &nshard) != 2 || nshard < 1 || shard < 0 || shard >= nshard) {
                cout << "Error: --shard expects i/N" << endl;
                return 1;
            }
        } else {
            cout << "usage: main_cohort_train --cfg exp_setup.json --pans list.txt [--device d] [--shard i/N] [--resume] [--host-optimizer] [--max-evals E]" << endl;
            return 1;
        }
    }
    if (exp_cfg.empty() || pan_file.empty()) {
        cout << "Error: --cfg and --pans are required" << endl;
        return 1;
    }
    c_experiment curr_exp(exp_cfg);
    const vector<int> kp = curr_exp.get_kernel_param();
    const int n_cov = curr_exp.get_cov_num(), n_lik = curr_exp.get_lik_num();

    // ---- load the cohort, deal shards by descending n^3
    vector<string> pans;
    {
        std::ifstream f(pan_file.c_str());
        string line;
        while (f >> line) pans.push_back(line);
    }
    // sizes-only pass over the whole cohort (one count line per feature file, over the host
    // cores), LPT deal, then only this shard's patients are parsed
    const double t_io = now_s();
    const vector<int> sizes = curr_exp.get_cohort_sizes(pans);
    const vector<int> shard_of = medgp_lpt_assign(sizes, nshard);
    vector<string> my_pans;
    size_t resumed = 0;
    for (size_t k = 0; k < pans.size(); k++) {
        if (shard_of[k] != shard) continue;
        if (resume) {  // already trained in an earlier run?
            std::ifstream flag((curr_exp.get_exp_train_dir() + "train_flag_" + pans[k] + ".txt").c_str());
            std::ifstream hyp((curr_exp.get_exp_train_dir() + "train_hyp_" + pans[k] + ".bin").c_str());
            int v = 0;
            if ((flag >> v) && v == 1 && hyp.good()) { resumed++; continue; }
        }
        my_pans.push_back(pans[k]);
    }
    if (resume) cout << "resume: " << resumed << " patients of this shard already have results" << endl;
    vector<c_experiment::patient_data> loaded;
    curr_exp.get_cohort_data(my_pans, loaded);
    vector<Patient> all(my_pans.size());
    vector<Patient *> mine;
    for (size_t k = 0; k < my_pans.size(); k++) {
        all[k].pan = my_pans[k];
        all[k].meta.swap(loaded[k].meta);
        all[k].time.swap(loaded[k].time);
        all[k].value.swap(loaded[k].value);
        mine.push_back(&all[k]);
    }
    // largest first: the library sorts evaluations by size anyway, and output order does not matter
    std::stable_sort(mine.begin(), mine.end(), [](const Patient *a, const Patient *b) { return a->time.size() > b->time.size(); });
    cout << "loaded " << mine.size() << " of " << pans.size() << " patients in " << now_s() - t_io << " s" << endl;
    cout << "shard " << shard << "/" << nshard << ": " << mine.size() << " of " << pans.size() << " patients on device " << device << endl;

    medgp_ctx *ctx = medgp_backend::context(kp[0], kp[1], kp[2], device);
    c_objective_batch batch(ctx, kp[0], kp[1], kp[2]);
    vector<vector<double> > global_hyp_array;
    curr_exp.get_global_hyp(global_hyp_array);  // identical for every patient, as in the reference
    const int n_init = curr_exp.get_scg_init_num();
    const double t0 = now_s();
    long total_evals = 0;

    // ---- data check + upload
    for (Patient *p : mine) {
        vector<int> count(curr_exp.get_feature_index().size(), 0);
        for (int m : p->meta) count[m]++;
        p->enough = !p->time.empty() && *std::min_element(count.begin(), count.end()) >= 2 && p->time.size() > 2;
        if (!p->enough) continue;
        if (medgp_cuda_add_series(ctx, (int)p->time.size(), (const int32_t *)p->meta.data(), p->time.data(),
                                  p->value.data(), &p->series_id) != MEDGP_OK) {
            std::cerr << "ERROR: " << medgp_cuda_last_error(ctx) << endl;
            return 1;
        }
        p->prior.reset(new c_prior(n_cov, 0, n_lik));
    }
    // ---- phase A: score the random initialisations (NLML only), a few patients per call
    {
        const size_t per_call = std::max<size_t>(1, 8192 / std::max(1, n_init));
        for (size_t b = 0; b < mine.size(); b += per_call) {
            vector<medgp_eval_request> reqs;
            vector<Patient *> owner;
            for (size_t k = b; k < std::min(mine.size(), b + per_call); k++) {
                if (!mine[k]->enough) continue;
                for (int r = 0; r < n_init; r++) {
                    reqs.push_back({mine[k]->series_id, &global_hyp_array[r], mine[k]->prior.get()});
                    owner.push_back(mine[k]);
                }
            }
            vector<medgp_eval_result> res;
            batch.compute(false, reqs, res);
            total_evals += (long)reqs.size();
            for (size_t q = 0; q < reqs.size(); q += n_init) {
                Patient *p = owner[q];
                for (int r = 0; r < n_init; r++) {  // the reference stops at the first failure
                    p->success = res[q + r].ok;
                    if (!p->success) break;
                    if (res[q + r].value < p->best_loss) {
                        p->best_loss = res[q + r].value;
                        p->best_init = global_hyp_array[r];
                    }
                }
            }
        }
    }
    cout << "phase A (random initialisations): " << total_evals << " NLML evaluations in " << now_s() - t0 << " s" << endl;
    // ---- phase B: optimise every patient's best initialisation in lock-step
    const double tB = now_s();
    for (Patient *p : mine) {
        if (!p->enough) continue;
        curr_exp.output_double_bin(curr_exp.get_exp_train_dir() + "train_init_hyp_" + p->pan, p->best_init);
        if (!p->success) continue;
        p->prior->setup_param(curr_exp.get_kernel_index(), kp, curr_exp.get_prior_mode(), curr_exp.get_prior_hyp());
        p->use_vem = curr_exp.get_prior_mode() == 2;
        if (p->use_vem)
            p->vem = varem_stepper(-curr_exp.get_scg_max_iter_num(), p->best_init, curr_exp.get_prior_sub_opt_iter(),
                                   kp, n_lik, p->prior.get());
        else
            p->scg = scg_stepper(-curr_exp.get_scg_max_iter_num(), p->best_init);
        p->active = true;
    }
    long grad_evals = 0, super_steps = 0;
    std::vector<medgp_opt_instance> inst;
    std::vector<Patient *> inst_owner;
    for (Patient *p : mine)
        if (p->active) {
            medgp_opt_instance it;
            it.series_id = p->series_id;
            it.init_parameter = p->best_init;
            it.prior = p->prior.get();
            it.use_varem = p->use_vem;
            it.max_iteration = -curr_exp.get_scg_max_iter_num();
            it.sub_opt_iter = curr_exp.get_prior_sub_opt_iter();
            inst.push_back(it);
            inst_owner.push_back(p);
        }
    const bool on_device = !host_optimizer && max_evals < 0 && medgp_device_optimizer_supports(inst);
    if (on_device) {
        super_steps = medgp_optimize_on_device(ctx, kp, n_lik, inst);
        for (size_t k = 0; k < inst.size(); k++) {
            inst_owner[k]->opt_parameter = inst[k].opt_parameter;
            grad_evals += inst[k].evals;
        }
    } else {
      while (true) {
        vector<medgp_eval_request> reqs;
        vector<Patient *> owner;
        for (Patient *p : mine)
            if (p->active && p->wants_eval()) {
                reqs.push_back({p->series_id, &p->point(), p->prior.get()});
                owner.push_back(p);
            }
        if (reqs.empty() || (max_evals >= 0 && grad_evals >= max_evals)) break;
        vector<medgp_eval_result> res;
        batch.compute(true, reqs, res);
        // every request of a super-step belongs to a different patient: the steppers advance independently
#pragma omp parallel for schedule(dynamic, 8)
        for (long q = 0; q < (long)reqs.size(); q++) owner[q]->feed(res[q].ok, res[q].value, res[q].grad);
        grad_evals += (long)reqs.size();
        super_steps++;
      }
    }
    const double dtB = now_s() - tB;
    cout << "phase B (optimisation, " << (on_device ? "device-resident" : "host") << " line searches): " << grad_evals << " NLML+gradient evaluations in " << super_steps
         << " super-steps, " << dtB << " s (" << (dtB > 0 ? grad_evals / dtB : 0.0) << " evals/s)" << endl;
    // ---- outputs (independent files per patient: written over the host cores)
#pragma omp parallel for schedule(dynamic, 8)
    for (long k = 0; k < (long)mine.size(); k++) {
        Patient *p = mine[k];
        bool flag = false;
        if (p->enough && p->success) {
            if (!on_device) p->opt_parameter = p->use_vem ? p->vem.best_parameter() : p->scg.best_parameter();
            curr_exp.output_double_bin(curr_exp.get_exp_train_dir() + "train_hyp_" + p->pan, p->opt_parameter);
            if (p->use_vem)
                curr_exp.output_double_bin(curr_exp.get_exp_train_dir() + "train_var_hyp_" + p->pan, p->prior->get_cov_varEM_all());
            flag = true;
        }
        curr_exp.output_int_txt(curr_exp.get_exp_train_dir() + "train_num_" + p->pan, vector<int>(1, (int)p->time.size()));
        curr_exp.output_int_txt(curr_exp.get_exp_train_dir() + "train_flag_" + p->pan, vector<int>(1, (int)flag));
    }
    total_evals += grad_evals;
    cout << "Finish all jobs. " << total_evals << " evaluations, total elapsed time = " << now_s() - t0 << " seconds" << endl;
    batch.release();
    medgp_backend::shutdown();
    return 0;
}
