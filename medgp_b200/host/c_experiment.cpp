// c_experiment.cpp -- see c_experiment.h.  File formats: SURVEY.md appendix B.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "c_experiment.h"
#include "medgp_host.h"

using std::string;
using std::vector;

namespace {
// ---- minimal JSON: enough for the flat exp_setup.json written by medgpc/util/config.py:5-35
struct JsonValue {
    enum Kind { NUL, NUM, STR, BOOL, OTHER } kind = NUL;
    double num = 0.0;
    bool is_int = false;
    string str;
};

struct JsonReader {
    const string &s;
    size_t p = 0;
    explicit JsonReader(const string &text) : s(text) {}
    void ws() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    [[noreturn]] void fail(const char *msg)
    {
        std::cerr << "ERROR: config JSON: " << msg << " at offset " << p << std::endl;
        exit(1);
    }
    string parse_string()
    {
        if (s[p] != '"') fail("expected string");
        p++;
        string out;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                p++;
                switch (s[p]) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'u': p += 4; out += '?'; break;
                    default: out += s[p];
                }
            } else {
                out += s[p];
            }
            p++;
        }
        if (p >= s.size()) fail("unterminated string");
        p++;
        return out;
    }
    void skip_compound(char open, char close)
    {
        int depth = 0;
        bool in_str = false;
        for (; p < s.size(); p++) {
            const char c = s[p];
            if (in_str) {
                if (c == '\\') p++;
                else if (c == '"') in_str = false;
            } else if (c == '"') in_str = true;
            else if (c == open) depth++;
            else if (c == close && --depth == 0) { p++; return; }
        }
        fail("unterminated array/object");
    }
    JsonValue parse_value()
    {
        ws();
        JsonValue v;
        if (p >= s.size()) fail("unexpected end");
        const char c = s[p];
        if (c == '"') { v.kind = JsonValue::STR; v.str = parse_string(); }
        else if (c == '{') { v.kind = JsonValue::OTHER; skip_compound('{', '}'); }
        else if (c == '[') { v.kind = JsonValue::OTHER; skip_compound('[', ']'); }
        else if (!s.compare(p, 4, "true")) { v.kind = JsonValue::BOOL; v.num = 1; p += 4; }
        else if (!s.compare(p, 5, "false")) { v.kind = JsonValue::BOOL; v.num = 0; p += 5; }
        else if (!s.compare(p, 4, "null")) { p += 4; }
        else {
            const size_t b = p;
            while (p < s.size() && (isdigit((unsigned char)s[p]) || strchr("+-.eE", s[p]))) p++;
            if (b == p) fail("unexpected character");
            const string tok = s.substr(b, p - b);
            v.kind = JsonValue::NUM;
            v.num = atof(tok.c_str());
            v.is_int = tok.find_first_of(".eE") == string::npos;
        }
        return v;
    }
    std::map<string, JsonValue> parse_object()
    {
        std::map<string, JsonValue> out;
        ws();
        if (p >= s.size() || s[p] != '{') fail("expected object");
        p++;
        ws();
        if (s[p] == '}') { p++; return out; }
        while (true) {
            ws();
            const string key = parse_string();
            ws();
            if (s[p] != ':') fail("expected ':'");
            p++;
            out[key] = parse_value();
            ws();
            if (s[p] == ',') { p++; continue; }
            if (s[p] == '}') { p++; break; }
            fail("expected ',' or '}'");
        }
        return out;
    }
};

const JsonValue &need(const std::map<string, JsonValue> &d, const char *key, JsonValue::Kind kind)
{
    auto it = d.find(key);
    if (it == d.end() || it->second.kind != kind) {
        std::cerr << "ERROR: config key \"" << key << "\" missing or of the wrong type" << std::endl;
        exit(1);
    }
    return it->second;
}
}  // namespace

c_experiment::c_experiment() { init_default_param(); }

void c_experiment::init_default_param()
{
    srand_seed = 718;
    kernel_index = 7;
    kernel_param.assign(3, 1);
    cv_fold_num = 1;
    scg_init_num = 1;
    scg_max_iter_num = 100;
    prior_mode = 0;
    prior_sub_opt_iter = DEFAULT_SCG_MAX_ITER;
    learn_rate = 1e-5;
    momentum = 0.9;
}

c_experiment::c_experiment(const string &input_cfg_name)
{
    init_default_param();
    exp_cfg_file = input_cfg_name;
    std::cout << "read in config. file " << exp_cfg_file << std::endl;
    std::ifstream ifs(exp_cfg_file.c_str());
    if (!ifs) {
        std::cerr << "File " << exp_cfg_file << " could not be opened." << std::endl;
        exit(1);
    }
    std::stringstream buf;
    buf << ifs.rdbuf();
    const string text = buf.str();
    JsonReader rd(text);
    const std::map<string, JsonValue> d = rd.parse_object();

    exp_data_dir = need(d, "data_dir", JsonValue::STR).str + "/";
    exp_top_dir = need(d, "exp_top_dir", JsonValue::STR).str + "/";
    exp_train_dir = need(d, "exp_train_dir", JsonValue::STR).str + "/";
    exp_test_dir = need(d, "exp_test_dir", JsonValue::STR).str + "/";
    exp_kernel_dir = need(d, "exp_kernel_dir", JsonValue::STR).str + "/";
    kernel_index = (int)need(d, "kernel_index", JsonValue::NUM).num;
    kernel_param.clear();
    kernel_param.push_back((int)need(d, "Q", JsonValue::NUM).num);
    kernel_param.push_back((int)need(d, "D", JsonValue::NUM).num);
    kernel_param.push_back((int)need(d, "R", JsonValue::NUM).num);
    for (int i = 0; i < 3; i++) std::cout << "kernel_param[" << i << "] = " << kernel_param[i] << std::endl;
    prior_mode = (int)need(d, "prior_index", JsonValue::NUM).num;
    if (prior_mode == 2) {
        prior_hyp.clear();
        prior_hyp.push_back((float)need(d, "eta", JsonValue::NUM).num);
        prior_hyp.push_back((float)need(d, "beta_lam", JsonValue::NUM).num);
    }
    std::istringstream is(need(d, "feature_index", JsonValue::STR).str);
    feature_index.clear();
    for (int k = 0; k < kernel_param[1]; k++) {
        int val = 0;
        is >> val;
        feature_index.push_back(val);
    }
    srand_seed = (int)need(d, "random_seed", JsonValue::NUM).num;
    cv_fold_num = (int)need(d, "cv_fold_num", JsonValue::NUM).num;
    scg_init_num = (int)need(d, "random_init_num", JsonValue::NUM).num;
    scg_max_iter_num = (int)need(d, "top_iteration_num", JsonValue::NUM).num;
    prior_sub_opt_iter = (int)need(d, "iteration_num_per_update", JsonValue::NUM).num;
    learn_rate = need(d, "online_learn_rate", JsonValue::NUM).num;
    momentum = need(d, "online_momentum", JsonValue::NUM).num;
    exp_hyp_bound_file = need(d, "exp_cfg_dir", JsonValue::STR).str + "/" + need(d, "hyp_bound_file", JsonValue::STR).str;
    get_hyp_bounds();
    print_experiment();
}

vector<int> c_experiment::get_lik_param() const
{
    vector<int> lik_param;
    if (get_lik_num() >= 1) lik_param.push_back(kernel_param[1]);
    return lik_param;
}

int c_experiment::get_cov_num() const
{
    const int Q = kernel_param[0], D = kernel_param[1], R = kernel_param[2];
    if (kernel_index != 7) {
        std::cout << "ERROR: the GPU backend supports kernel_index 7 (LMC-SM) only, got " << kernel_index << std::endl;
        exit(1);
    }
    return Q * (D * R + 2 + D);
}

int c_experiment::get_lik_num() const { return kernel_param[1]; }

void c_experiment::get_one_patient_data(string PAN, vector<int> &meta_vec, vector<float> &time_vec,
                                        vector<float> &value_vec, bool verbose) const
{
    meta_vec.clear();
    time_vec.clear();
    value_vec.clear();
    for (int j = 0; j < (int)feature_index.size(); j++) {
        const string fid = std::to_string((long long)feature_index[j]);
        // cohort mean / standard deviation: two raw doubles
        vector<double> stat;
        {
            std::ifstream databin(exp_data_dir + "feature" + fid + "_stat.bin", std::ios::binary);
            double f;
            while (databin.read(reinterpret_cast<char *>(&f), sizeof(double))) stat.push_back(f);
        }
        if (stat.size() < 2) {
            std::cerr << "File " << exp_data_dir << "feature" << fid << "_stat.bin is missing or short." << std::endl;
            exit(1);
        }
        const string filename = exp_data_dir + PAN + "/feature" + fid + ".txt";
        std::ifstream data(filename.c_str());
        if (!data) {
            std::cerr << "File " << filename << " could not be opened." << std::endl;
            exit(1);
        }
        if (verbose) std::cout << "reading data file " << filename << " (mean/std " << stat[0] << " " << stat[1] << ")" << std::endl;
        float vec_len = 0, temp = 0;
        data >> vec_len;
        for (int i = 0; i < (int)vec_len; i++) {
            meta_vec.push_back(j);
            data >> temp;
            time_vec.push_back(temp);
            data >> temp;
            const double norm_temp = ((double)temp - stat[0]) / stat[1];
            value_vec.push_back((float)norm_temp);
        }
    }
}

namespace {
// [mean, std] of one feature over the cohort: two raw doubles (c_experiment.cpp:276-284)
vector<double> read_feature_stat(const string &data_dir, const string &fid)
{
    vector<double> stat;
    std::ifstream databin(data_dir + "feature" + fid + "_stat.bin", std::ios::binary);
    double f;
    while (databin.read(reinterpret_cast<char *>(&f), sizeof(double))) stat.push_back(f);
    if (stat.size() < 2) {
        std::cerr << "File " << data_dir << "feature" << fid << "_stat.bin is missing or short." << std::endl;
        exit(1);
    }
    return stat;
}
}  // namespace

vector<int> c_experiment::get_cohort_sizes(const vector<string> &pans) const
{
    vector<int> sizes(pans.size(), 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (long k = 0; k < (long)pans.size(); k++) {
        int n = 0;
        for (int j = 0; j < (int)feature_index.size(); j++) {
            const string filename = exp_data_dir + pans[k] + "/feature" + std::to_string((long long)feature_index[j]) + ".txt";
            std::ifstream data(filename.c_str());
            if (!data) {
#pragma omp critical
                std::cerr << "File " << filename << " could not be opened." << std::endl;
                exit(1);
            }
            float vec_len = 0;
            data >> vec_len;  // the count line
            n += (int)vec_len;
        }
        sizes[k] = n;
    }
    return sizes;
}

void c_experiment::get_cohort_data(const vector<string> &pans, vector<patient_data> &out) const
{
    vector<vector<double> > stat(feature_index.size());
    for (size_t j = 0; j < feature_index.size(); j++)
        stat[j] = read_feature_stat(exp_data_dir, std::to_string((long long)feature_index[j]));
    out.assign(pans.size(), patient_data());
#pragma omp parallel for schedule(dynamic, 4)
    for (long k = 0; k < (long)pans.size(); k++) {
        patient_data &p = out[k];
        for (int j = 0; j < (int)feature_index.size(); j++) {
            const string filename = exp_data_dir + pans[k] + "/feature" + std::to_string((long long)feature_index[j]) + ".txt";
            std::ifstream data(filename.c_str());
            if (!data) {
#pragma omp critical
                std::cerr << "File " << filename << " could not be opened." << std::endl;
                exit(1);
            }
            float vec_len = 0, temp = 0;
            data >> vec_len;
            for (int i = 0; i < (int)vec_len; i++) {  // same arithmetic as get_one_patient_data
                p.meta.push_back(j);
                data >> temp;
                p.time.push_back(temp);
                data >> temp;
                const double norm_temp = ((double)temp - stat[j][0]) / stat[j][1];
                p.value.push_back((float)norm_temp);
            }
        }
    }
}

vector<int> medgp_lpt_assign(const vector<int> &sizes, int nshard)
{
    vector<size_t> order(sizes.size());
    for (size_t k = 0; k < order.size(); k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return sizes[a] > sizes[b]; });
    vector<double> load(std::max(1, nshard), 0.0);
    vector<int> out(sizes.size(), 0);
    for (size_t k : order) {
        const int tgt = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        const double n = (double)sizes[k];
        load[tgt] += n * n * n;
        out[k] = tgt;
    }
    return out;
}

void c_experiment::get_hyp_bounds()
{
    std::ifstream data(exp_hyp_bound_file.c_str());
    if (!data) {
        std::cerr << "File " << exp_hyp_bound_file << " could not be opened." << std::endl;
        exit(1);
    }
    hyp_array_lb.clear();
    hyp_array_ub.clear();
    double temp;
    for (int i = 0; i < get_hyp_num(); i++) {
        data >> temp;
        hyp_array_lb.push_back(temp);
        data >> temp;
        hyp_array_ub.push_back(temp);
    }
}

// one rand() per hyper-parameter, hyper-major within an init, init-major overall, after a
// single srand(seed): bit-identical draws to the reference (c_experiment.cpp:418-441,493-564)
void c_experiment::get_global_hyp(vector<vector<double> > &global_hyp_array)
{
    std::cout << "generating random hyperparameters..." << std::endl;
    srand(srand_seed);
    for (int i = 0; i < scg_init_num; i++) {
        vector<double> hyp_array;
        get_hyp_LMC_SM(hyp_array);
        global_hyp_array.push_back(hyp_array);
    }
}

double c_experiment::get_one_random(const double &lb, const double &ub, const double &scale,
                                    const bool &flag_inv, const bool &flag_log)
{
    const int rand_max = (int)floor(pow(2.0, 12));
    double temp = ((double)(rand() % rand_max)) + 1.0;
    temp *= (ub - lb);
    temp = temp / ((double)rand_max);
    double a = scale * (temp + lb);
    if (flag_inv) a = 1.0 / a;
    if (flag_log) a = log(a);
    return a;
}

void c_experiment::get_hyp_LMC_SM(vector<double> &hyp_array)
{
    const int Q = kernel_param[0], D = kernel_param[1], R = kernel_param[2];
    const int nl = get_lik_num();
    for (int i = 0; i < get_hyp_num(); i++) {
        double temp;
        if (i < nl) {  // noise
            temp = get_one_random(hyp_array_lb[i], hyp_array_ub[i], 1.0, false, true);
        } else if (i < nl + Q * D * R) {  // A, raw
            const double dQ = Q, dR = R;
            temp = get_one_random(hyp_array_lb[i], hyp_array_ub[i], 0.9 / sqrt(dQ * dR), false, false);
        } else if (i < nl + Q * (D * R + 1)) {  // mu = 1 / period
            temp = get_one_random(hyp_array_lb[i], hyp_array_ub[i], 1.0, false, false);
            temp = log(1.0 / temp);
        } else if (i < nl + Q * (D * R + 2)) {  // v = 1 / (2 PI lengthscale)
            temp = get_one_random(hyp_array_lb[i], hyp_array_ub[i], 1.0, false, false);
            temp = log(1.0 / (2 * PI * temp));
        } else {  // kappa
            const double dQ = Q;
            temp = get_one_random(hyp_array_lb[i], hyp_array_ub[i], 0.1 / dQ, false, true);
        }
        hyp_array.push_back(temp);
    }
}

void c_experiment::print_experiment() const
{
    std::cout << "---------------------------------------------------" << std::endl
              << "Summary of the experiment:" << std::endl
              << "Input config. file: " << exp_cfg_file << std::endl
              << "Input data path: " << exp_data_dir << std::endl
              << "Training result output path: " << exp_train_dir << std::endl
              << "Testing result output path: " << exp_test_dir << std::endl
              << "Index of testing feature(s): ";
    for (size_t i = 0; i < feature_index.size(); i++) std::cout << feature_index[i] << ' ';
    std::cout << std::endl
              << "Total # of CV fold: " << cv_fold_num << std::endl
              << "Current kernel index: " << kernel_index << std::endl
              << "Current random seed: " << srand_seed << std::endl
              << "Loading hyperparameters boundary from: " << exp_hyp_bound_file << std::endl
              << "Total number of hyperparameters: " << get_hyp_num() << std::endl
              << "---------------------------------------------------" << std::endl;
}

void c_experiment::output_double_bin(const string &file_prefix, vector<double> hyp_array) const
{
    std::ofstream data((file_prefix + ".bin").c_str(), std::ios::binary);
    if (data.is_open())
        data.write(reinterpret_cast<const char *>(hyp_array.data()), sizeof(double) * hyp_array.size());
}

void c_experiment::output_int_txt(const string &file_prefix, vector<int> int_array) const
{
    std::ofstream data((file_prefix + ".txt").c_str());
    if (data.is_open())
        for (size_t i = 0; i < int_array.size(); i++) data << int_array[i] << "\n";
}

vector<int> c_experiment::get_test_kernel_param(int fold, const string &kernel_clust_alg) const
{
    vector<int> test_kernel_param(kernel_param);
    const string f = exp_kernel_dir + "fold" + std::to_string((long long)fold) + "/" + kernel_clust_alg +
                     "_mode_mixture_num.txt";
    std::ifstream data(f.c_str());
    int q = 0;
    if (!(data >> q)) {
        std::cerr << "File " << f << " could not be read." << std::endl;
        exit(1);
    }
    test_kernel_param[0] = q;
    return test_kernel_param;
}

int c_experiment::get_test_cov_num(int fold, const string &kernel_clust_alg) const
{
    const vector<int> kp = get_test_kernel_param(fold, kernel_clust_alg);
    return kp[0] * (kp[1] * kp[2] + 2 + kp[1]);
}

vector<double> c_experiment::get_test_mode_param(int fold, const string &kernel_clust_alg) const
{
    vector<double> mode_param;
    const string f = exp_kernel_dir + "fold" + std::to_string((long long)fold) + "/" + kernel_clust_alg +
                     "_mode_param.bin";
    std::ifstream databin(f, std::ios::binary);
    double one;
    while (databin.read(reinterpret_cast<char *>(&one), sizeof(double))) mode_param.push_back(one);
    std::cout << "read in " << mode_param.size() << " mode parameters from " << f << std::endl;
    return mode_param;
}
