// optimizer.cpp -- re-entrant SCG (Rasmussen minimize) and variational-EM steppers plus the
// blocking c_optimizer_scg / c_optimizer_varEM built on them.  Evaluation-for-evaluation the
// same control flow as medgpc/src/util/c_optimizer_scg.cpp:25-284 and
// c_optimizer_varEM.cpp:26-163, including the reference's quirks (see comments).
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>

#include "medgp_host.h"

using std::vector;

namespace {
const double INT_ = 0.1, EXT = 3.0, MAXEV = 20, RATIO = 10, SIG = 0.1, RHO = SIG / 2.0;

double dot(const vector<double> &a, const vector<double> &b)
{
    double s = 0.0;
    for (size_t k = 0; k < a.size(); k++) s += a[k] * b[k];
    return s;
}
inline int neg_flag(int length) { return std::signbit((double)length) ? 1 : 0; }
}  // namespace

// ------------------------------------------------------------------ scg_stepper
scg_stepper::scg_stepper(int max_iteration, const vector<double> &init_parameter)
    : state(INIT), length(max_iteration), i(0), n_eval(0), ls_failed(false), obj_flag(false),
      success(false), M(0), d0(0), f0_unused(0), x1(0), x2(0), x3(0), x4(0), d1(0), d2(0), d3(0), d4(0),
      f1(0), f2(0), f3(0), f4(0), F0(0), fX(0), X(init_parameter), probe(init_parameter)
{
    // The reference advances its counter by signbit(max_iteration) only (c_optimizer_scg.cpp:73,88,114),
    // so a positive budget ("Linesearch" mode) never terminates there.  Nothing in MedGP passes one;
    // it is refused here instead of being reproduced.
    if (max_iteration > 0) {
        std::cout << "ERROR: c_optimizer_scg needs a negative max_iteration (function-evaluation budget); got "
                  << max_iteration << std::endl;
        fX = NAN;
        state = DONE;
    }
}

void scg_stepper::make_probe()
{
    probe.resize(X.size());
    for (size_t k = 0; k < X.size(); k++) probe[k] = X[k] + x3 * s[k];
}

void scg_stepper::feed(bool ok, double f, const vector<double> &g)
{
    n_eval++;
    if (state == INIT) {
        if (!ok) {  // the reference would continue with indeterminate values; stop instead
            fX = NAN;
            state = DONE;
            return;
        }
        fX = f;
        df0 = g;
        i = i + neg_flag(length);
        s.resize(df0.size());
        for (size_t k = 0; k < df0.size(); k++) s[k] = -1.0 * df0[k];
        d0 = -1.0 * dot(s, s);
        x3 = 1.0 / (1.0 - d0);  // red = 1
        begin_iteration();
        return;
    }
    if (state == EXTRAPOLATE) {
        obj_flag = ok;
        if (ok) { f3 = f; df3 = g; }
        if (!ok || std::isinf(f3) || std::isnan(f3)) x3 = (x2 + x3) / 2.0;
        else success = true;
        request_extrapolation_eval();
        return;
    }
    if (state == INTERPOLATE) {
        obj_flag = ok;
        if (ok) { f3 = f; df3 = g; }
        if (obj_flag && f3 < F0) {
            X0 = probe;
            F0 = f3;
            dF0 = df3;
        }
        M = M - 1;
        i = i + neg_flag(length);
        d3 = dot(df3, s);
        continue_interpolation();
        return;
    }
}

void scg_stepper::begin_iteration()
{
    if (!(i < std::abs(length))) {
        state = DONE;
        return;
    }
    i = i + neg_flag(length);  // reference: signbit here as well (c_optimizer_scg.cpp:88)
    X0 = X;
    F0 = fX;
    dF0 = df0;
    M = (length > 0) ? MAXEV : std::min((int)MAXEV, std::abs(length) - i);
    begin_extrapolation_pass();
}

void scg_stepper::begin_extrapolation_pass()
{
    // the reference re-initialises these at the top of EVERY pass of its while(1) loop
    // (c_optimizer_scg.cpp:101-110), so x1 is always 0 in the cubic extrapolation below
    x2 = 0.0;
    f2 = fX;
    d2 = d0;
    f3 = fX;
    df3 = df0;
    success = false;
    request_extrapolation_eval();
}

void scg_stepper::request_extrapolation_eval()
{
    if (!success && M > 0) {
        M = M - 1;
        i = i + neg_flag(length);
        make_probe();
        state = EXTRAPOLATE;
        return;
    }
    after_extrapolation_eval();
}

void scg_stepper::after_extrapolation_eval()
{
    if (f3 < F0) {
        X0.resize(X.size());
        for (size_t k = 0; k < X.size(); k++) X0[k] = X[k] + x3 * s[k];
        F0 = f3;
        dF0 = df3;
    }
    d3 = dot(df3, s);
    if ((d3 > SIG * d0) || (f3 > (fX + x3 * RHO * d0)) || (M == 0)) {
        continue_interpolation();
        return;
    }
    x1 = x2; f1 = f2; d1 = d2;
    x2 = x3; f2 = f3; d2 = d3;
    const double A = 6.0 * (f1 - f2) + 3.0 * (d2 + d1) * (x2 - x1);
    const double B = 3.0 * (f2 - f1) - (2.0 * d1 + d2) * (x2 - x1);
    const double temp = B * B - A * d1 * (x2 - x1);
    if (temp < 0) {
        x3 = x2 * EXT;
    } else {
        x3 = x1 - (d1 * pow(x2 - x1, 2.0) / (B + sqrt(temp)));
        if (std::isnan(x3) || std::isinf(x3) || (x3 < 0)) x3 = x2 * EXT;
        else if (x3 > x2 * EXT) x3 = x2 * EXT;
        else if (x3 < (x2 + INT_ * (x2 - x1))) x3 = x2 + INT_ * (x2 - x1);
    }
    begin_extrapolation_pass();
}

void scg_stepper::continue_interpolation()
{
    if (((fabs(d3) > -1.0 * SIG * d0) || (f3 > (fX + x3 * RHO * d0))) && (M > 0)) {
        if ((d3 > 0) || (f3 > (fX + x3 * RHO * d0))) { x4 = x3; f4 = f3; d4 = d3; }
        else { x2 = x3; f2 = f3; d2 = d3; }
        if (f4 > fX) {
            x3 = x2 - (0.5 * d2 * pow(x4 - x2, 2.0)) / (f4 - f2 - d2 * (x4 - x2));
            if (std::isnan(x3) || std::isinf(x3)) x3 = (x2 + x4) / 2.0;
        } else {
            const double A = 6.0 * (f2 - f4) / (x4 - x2) + 3.0 * (d4 + d2);
            const double B = 3.0 * (f4 - f2) - (2.0 * d2 + d4) * (x4 - x2);
            const double disc = B * B - A * d2 * pow(x4 - x2, 2.0);
            if (disc < 0) {
                x3 = (x2 + x4) / 2.0;
            } else {
                x3 = x2 + (sqrt(disc) - B) / A;
                if (std::isnan(x3) || std::isinf(x3)) x3 = (x2 + x4) / 2.0;
            }
        }
        x3 = std::max(std::min(x3, x4 - INT_ * (x4 - x2)), x2 + INT_ * (x4 - x2));
        make_probe();
        state = INTERPOLATE;
        return;
    }
    finish_iteration();
}

void scg_stepper::finish_iteration()
{
    if (obj_flag && (fabs(d3) < -1.0 * SIG * d0) && (f3 < (fX + x3 * RHO * d0))) {
        for (size_t k = 0; k < X.size(); k++) X[k] = X[k] + x3 * s[k];
        fX = f3;
        // Polak-Ribiere direction
        const double g33 = dot(df3, df3), g30 = dot(df3, df0), g00 = dot(df0, df0);
        for (size_t k = 0; k < s.size(); k++) s[k] = ((g33 - g30) / g00) * s[k] - df3[k];
        df0 = df3;
        d3 = d0;
        d0 = dot(df0, s);
        if (d0 > 0) {
            for (size_t k = 0; k < s.size(); k++) s[k] = -1.0 * df0[k];
            d0 = -1.0 * dot(s, s);
        }
        x3 = x3 * std::min(RATIO, d3 / (d0 - pow(2.0, -52)));
        ls_failed = false;
    } else {
        X = X0;
        fX = F0;
        df0 = dF0;
        for (size_t k = 0; k < s.size(); k++) s[k] = -1.0 * df0[k];
        d0 = -1.0 * dot(s, s);
        x3 = 1.0 / (1.0 - d0);
        ls_failed = true;
    }
    begin_iteration();
}

// ------------------------------------------------------------------ blocking optimisers
void c_optimizer::print_optimizer() const { std::cout << "current optimizer: " << optimizer_name << std::endl; }

void c_optimizer_scg::optimize(const int &max_iteration, const vector<double> &init_parameter,
                               c_objective *objfunc, const bool &display, double &opt_loss,
                               vector<double> &opt_parameter, c_kernel *&input_kernel,
                               c_meanfunc *&input_meanfunc, c_likelihood *&input_likfunc,
                               c_inference *&input_inffunc, c_prior *&input_prior)
{
    scg_stepper st(max_iteration, init_parameter);
    const char *method = max_iteration > 0 ? "Linesearch " : "Function evaluation ";
    double last = NAN;
    while (st.wants_eval()) {
        double f = 0.0;
        vector<double> g;
        const bool ok = objfunc->compute_objective(true, st.point(), f, g, input_kernel, input_meanfunc,
                                                   input_likfunc, input_inffunc, input_prior);
        st.feed(ok, f, g);
        if (display && st.best_loss() != last) {
            last = st.best_loss();
            std::cout << method << st.evaluations() << ": " << last << std::endl;
        }
    }
    opt_loss = st.best_loss();
    opt_parameter = st.best_parameter();
}

// ------------------------------------------------------------------ variational EM
double c_optimizer_varEM::update_tau(const float &gamma, const float &d, const float &eta, const double &phi)
{
    return (gamma + d) / (phi + eta);
}
double c_optimizer_varEM::update_phi(const int &D, const float &beta, const float &gamma,
                                     const double &delta_sum, const double &tau)
{
    return (((float)(D)) * beta + gamma - 1.0) / (delta_sum + tau);
}
double c_optimizer_varEM::update_delta(const float &alpha, const float &beta, const double &psi,
                                       const double &phi)
{
    return (alpha + beta) / (psi + phi);
}
double c_optimizer_varEM::update_psi(const float &alpha, const double &a, const double &delta)
{
    const double sub = (2.0 * alpha - 3.0);
    return (sub + sqrt(sub * sub + 8.0 * delta * a * a)) / (4.0 * delta);
}

varem_rounds::varem_rounds(int max_iteration, const vector<double> &init_parameter, int sub_iter,
                           const vector<int> &kernel_param, int lik_num_, c_prior *prior_)
    : prior(prior_), done_(false), max_iter(std::abs(max_iteration)), iter(0), sub_opt_iter(sub_iter),
      lik_num(lik_num_), opt_loss(0.0), best_loss_(0.0), opt_parameter(init_parameter)
{
    if (kernel_param.size() != 3) {
        std::cout << "ERROR: varEM is only usable for LMCSM kernel!" << std::endl;
        exit(1);
    }
    Q = kernel_param[0]; D = kernel_param[1]; R = kernel_param[2];
    if (max_iter == 0) done_ = true;
}

void varem_rounds::finish_round(double loss, const vector<double> &parameter)
{
    opt_loss = loss;
    opt_parameter = parameter;
    if (iter > 0) {
        const double change_ratio = (opt_loss - best_loss_) / best_loss_;
        if (fabs(change_ratio) < 0.005) {  // early stop (c_optimizer_varEM.cpp:89-95)
            done_ = true;
            return;
        }
    }
    best_loss_ = opt_loss;
    const float alpha = prior->get_cov_varEM_fix_one(0), beta = prior->get_cov_varEM_fix_one(1);
    const float gamma = prior->get_cov_varEM_fix_one(2), dd = prior->get_cov_varEM_fix_one(3);
    const float eta = prior->get_cov_varEM_fix_one(4);
    // state layout: psi [0, QDR) | delta [QDR, 2QDR) | phi [2QDR, 2QDR+QR) | tau [2QDR+QR, ...)
    for (int q = 0; q < Q; q++)
        for (int r = 0; r < R; r++) {  // tau
            const int index = Q * (2 * D * R + R) + q * R + r;
            const double phi = prior->get_cov_varEM_one(index - Q * R);
            prior->set_cov_varEM_one(c_optimizer_varEM::update_tau(gamma, dd, eta, phi), index);
        }
    for (int q = 0; q < Q; q++)
        for (int r = 0; r < R; r++) {  // phi
            const int index = Q * (2 * D * R) + q * R + r;
            double delta_sum = 0.0;
            for (int d = 0; d < D; d++) delta_sum += prior->get_cov_varEM_one(Q * D * R + q * D * R + d * R + r);
            const double tau = prior->get_cov_varEM_one(index + Q * R);
            prior->set_cov_varEM_one(c_optimizer_varEM::update_phi(D, beta, gamma, delta_sum, tau), index);
        }
    for (int q = 0; q < Q; q++)
        for (int d = 0; d < D; d++)
            for (int r = 0; r < R; r++) {  // delta
                const int index = Q * D * R + q * D * R + d * R + r;
                const double psi = prior->get_cov_varEM_one(index - Q * D * R);
                const double phi = prior->get_cov_varEM_one(2 * Q * D * R + q * R + r);
                prior->set_cov_varEM_one(c_optimizer_varEM::update_delta(alpha, beta, psi, phi), index);
            }
    for (int q = 0; q < Q; q++)
        for (int d = 0; d < D; d++)
            for (int r = 0; r < R; r++) {  // psi, and pruning of A entries whose psi hits 0
                const int index = q * D * R + d * R + r;
                const double a = opt_parameter[lik_num + index];
                const double delta = prior->get_cov_varEM_one(index + Q * D * R);
                prior->set_cov_varEM_one(c_optimizer_varEM::update_psi(alpha, a, delta), index);
                if (prior->get_cov_varEM_one(index) == 0.0) {
                    prior->type_cov[index] = 0;
                    opt_parameter[lik_num + index] = 0.0;
                }
                prior->fix_param_cov[index][0] = 0;
                prior->fix_param_cov[index][1] = prior->get_cov_varEM_one(index);
            }
    iter++;
    if (iter >= max_iter) done_ = true;
}

varem_stepper::varem_stepper(int max_iteration, const vector<double> &init_parameter, int sub_iter,
                             const vector<int> &kernel_param, int lik_num, c_prior *prior)
    : outer(max_iteration, init_parameter, sub_iter, kernel_param, lik_num, prior)
{
    if (!outer.done()) scg = scg_stepper(outer.budget(), outer.start());
}

void varem_stepper::feed(bool ok, double f, const vector<double> &g)
{
    scg.feed(ok, f, g);
    if (scg.wants_eval()) return;
    outer.finish_round(scg.best_loss(), scg.best_parameter());
    if (!outer.done()) scg = scg_stepper(outer.budget(), outer.start());
}

// ------------------------------------------------------------------ device-resident lock-step optimisation
namespace {
// prior table of one instance in theta order [lik | cov] as the session wants it
bool prior_table(const c_prior *prior, int n_lik, int P, signed char *type, signed char *is_exp, float *param)
{
    for (int k = 0; k < P; k++) {
        type[k] = -1; is_exp[k] = 0; param[2 * k] = 0.f; param[2 * k + 1] = 1.f;
    }
    if (!prior) return true;
    for (int k = 0; k < P; k++) {
        const bool lik = k < n_lik;
        const int i = lik ? k : k - n_lik;
        const bool flag = lik ? prior->flag_lik[i] : prior->flag_cov[i];
        if (!flag) continue;
        const int t = lik ? prior->type_lik[i] : prior->type_cov[i];
        if (t > 2) return false;  // kde: not on the device
        // an active entry of type -1 changes nothing (c_inference_prior.cpp:106-108; one_lik returns zeros)
        type[k] = (signed char)(t < 0 ? -1 : t);
        is_exp[k] = (lik ? prior->exp_lik[i] : prior->exp_cov[i]) ? 1 : 0;
        const std::vector<float> &fp = lik ? prior->fix_param_lik[i] : prior->fix_param_cov[i];
        if (t >= 1 && fp.size() < 2) return false;
        if (fp.size() >= 2) { param[2 * k] = fp[0]; param[2 * k + 1] = fp[1]; }
    }
    return true;
}

void die_opt(const char *what, medgp_ctx *ctx, int rc)
{
    std::cerr << "ERROR: " << what << " failed with status " << rc << " (" << medgp_cuda_last_error(ctx)
              << "); libmedgp_cuda.so has no CPU fallback" << std::endl;
    exit(1);
}
}  // namespace

bool medgp_device_optimizer_supports(const std::vector<medgp_opt_instance> &inst)
{
    for (const auto &it : inst) {
        if (!it.prior) continue;
        for (size_t i = 0; i < it.prior->type_lik.size(); i++)
            if (it.prior->flag_lik[i] && it.prior->type_lik[i] > 2) return false;
        for (size_t i = 0; i < it.prior->type_cov.size(); i++)
            if (it.prior->flag_cov[i] && it.prior->type_cov[i] > 2) return false;
    }
    return true;
}

long medgp_optimize_on_device(medgp_ctx *ctx, const vector<int> &kernel_param, int lik_num,
                              std::vector<medgp_opt_instance> &inst, int poll_every,
                              medgp_external_objective external, void *user)
{
    const int P = medgp_cuda_num_hyp(ctx);
    long super_steps = 0;
    // outer state: plain SCG = one "round" with the whole budget
    std::vector<varem_rounds> outer(inst.size());
    std::vector<char> live(inst.size(), 0);
    for (size_t k = 0; k < inst.size(); k++) {
        medgp_opt_instance &it = inst[k];
        it.opt_parameter = it.init_parameter;
        it.opt_loss = NAN;
        it.evals = 0;
        if (it.use_varem) {
            outer[k] = varem_rounds(it.max_iteration, it.init_parameter, it.sub_opt_iter, kernel_param, lik_num, it.prior);
            live[k] = outer[k].done() ? 0 : 1;
        } else {
            live[k] = it.max_iteration < 0 ? 1 : 0;
            if (it.max_iteration > 0)
                std::cout << "ERROR: c_optimizer_scg needs a negative max_iteration (function-evaluation budget); got "
                          << it.max_iteration << std::endl;
        }
    }
    while (true) {
        // ---- the instances of this round
        std::vector<int> who;
        for (size_t k = 0; k < inst.size(); k++)
            if (live[k]) who.push_back((int)k);
        if (who.empty()) break;
        const int count = (int)who.size();
        std::vector<int> sids(count), budget(count);
        std::vector<double> theta0((size_t)count * P);
        std::vector<signed char> ptype((size_t)count * P), pexp((size_t)count * P);
        std::vector<float> ppar((size_t)count * P * 2);
        bool any_prior = false;
        for (int b = 0; b < count; b++) {
            medgp_opt_instance &it = inst[who[b]];
            sids[b] = it.series_id;
            budget[b] = it.use_varem ? outer[who[b]].budget() : it.max_iteration;
            const vector<double> &x0 = it.use_varem ? outer[who[b]].start() : it.init_parameter;
            std::copy(x0.begin(), x0.end(), theta0.begin() + (size_t)b * P);
            if (!prior_table(it.prior, lik_num, P, &ptype[(size_t)b * P], &pexp[(size_t)b * P], &ppar[(size_t)b * P * 2])) {
                std::cerr << "ERROR: this prior table cannot run on the device (kde prior); use the host optimisers" << std::endl;
                exit(1);
            }
            // an external objective (tests) owns its prior terms, as a c_objective does in the reference
            any_prior = any_prior || (it.prior != nullptr && !external);
        }
        medgp_scg *scg = nullptr;
        const bool trace = getenv("MEDGP_SCG_TRACE") != nullptr;
        auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double tr0 = tnow();
        int rc = medgp_cuda_scg_create(ctx, count, &scg);
        const double tr1 = tnow();
        if (rc != MEDGP_OK) die_opt("medgp_cuda_scg_create", ctx, rc);
        rc = medgp_cuda_scg_start(scg, sids.data(), theta0.data(), budget.data(), any_prior ? ptype.data() : nullptr,
                                  any_prior ? pexp.data() : nullptr, any_prior ? ppar.data() : nullptr);
        if (rc != MEDGP_OK) die_opt("medgp_cuda_scg_start", ctx, rc);
        const double tr2 = tnow();
        // ---- run the round to completion
        if (!external) {
            int left = count;
            while (left > 0) {
                rc = medgp_cuda_scg_run(scg, poll_every, &left);
                if (rc != MEDGP_OK) die_opt("medgp_cuda_scg_run", ctx, rc);
                super_steps += poll_every;
            }
        } else {  // tests: the same device state machine on an external objective
            std::vector<double> pts((size_t)count * P), f(count), g((size_t)count * P);
            std::vector<int> wants(count), ok(count);
            while (true) {
                rc = medgp_cuda_scg_points(scg, pts.data(), wants.data());
                if (rc != MEDGP_OK) die_opt("medgp_cuda_scg_points", ctx, rc);
                bool any = false;
                for (int b = 0; b < count; b++) {
                    ok[b] = 0;
                    if (!wants[b]) continue;
                    any = true;
                    vector<double> x(pts.begin() + (size_t)b * P, pts.begin() + (size_t)(b + 1) * P), gg;
                    double ff = 0.0;
                    ok[b] = external(who[b], x, ff, gg, user) ? 1 : 0;
                    if (ok[b]) {
                        f[b] = ff;
                        std::copy(gg.begin(), gg.end(), g.begin() + (size_t)b * P);
                    }
                }
                if (!any) break;
                rc = medgp_cuda_scg_feed(scg, f.data(), g.data(), ok.data());
                if (rc != MEDGP_OK) die_opt("medgp_cuda_scg_feed", ctx, rc);
                super_steps++;
            }
        }
        // ---- results, EM updates
        std::vector<double> best((size_t)count * P), loss(count);
        std::vector<int> evals(count);
        rc = medgp_cuda_scg_result(scg, best.data(), loss.data(), evals.data());
        if (rc != MEDGP_OK) die_opt("medgp_cuda_scg_result", ctx, rc);
        const double tr4 = tnow();
        medgp_cuda_scg_destroy(scg);
        if (trace)
            std::cerr << "scg round: count=" << count << " create " << tr1 - tr0 << " ms, start " << tr2 - tr1 << " ms, run+result "
                      << tr4 - tr2 << " ms, destroy " << tnow() - tr4 << " ms" << std::endl;
        for (int b = 0; b < count; b++) {
            medgp_opt_instance &it = inst[who[b]];
            const vector<double> x(best.begin() + (size_t)b * P, best.begin() + (size_t)(b + 1) * P);
            it.evals += evals[b];
            if (it.use_varem) {
                varem_rounds &o = outer[who[b]];
                o.finish_round(loss[b], x);
                it.opt_parameter = o.best_parameter();
                it.opt_loss = o.best_loss();
                live[who[b]] = o.done() ? 0 : 1;
            } else {
                it.opt_parameter = x;
                it.opt_loss = loss[b];
                live[who[b]] = 0;
            }
        }
    }
    return super_steps;
}

void c_optimizer_varEM::optimize(const int &max_iteration, const vector<double> &init_parameter,
                                 c_objective *objfunc, const bool &display, double &opt_loss,
                                 vector<double> &opt_parameter, c_kernel *&input_kernel,
                                 c_meanfunc *&input_meanfunc, c_likelihood *&input_likfunc,
                                 c_inference *&input_inffunc, c_prior *&input_prior)
{
    varem_stepper st(max_iteration, init_parameter, sub_opt_iter, input_kernel->get_kernel_param(),
                     input_likfunc->get_likfunc_hyp_num(), input_prior);
    int shown = 0;
    while (st.wants_eval()) {
        double f = 0.0;
        vector<double> g;
        const bool ok = objfunc->compute_objective(true, st.point(), f, g, input_kernel, input_meanfunc,
                                                   input_likfunc, input_inffunc, input_prior);
        st.feed(ok, f, g);
        if (display && st.rounds() != shown) {
            std::cout << "iteration " << shown << " for variational EM: loss = " << st.best_loss() << std::endl;
            shown = st.rounds();
        }
    }
    opt_loss = st.best_loss();
    opt_parameter = st.best_parameter();
}
