// ParameterReader.cpp -- see ParameterReader.h.  Keys and types follow reference
// src/ParameterReader.cpp:28-66 (27 typed fields + 5 camera globals); numeric values are re-formatted through
// a stringstream like the reference's num2string (so "0.080" reads back as "0.08").
// New OPTIONAL keys of the ICP backend (defaults keep the stock parameters.yaml loadable):
//   icp_iterations (10)  icp_max_corr_dist (0.2 m; 0 = unlimited like PCL's default)  icp_estimator (plane|svd)  icp_search (grid|brute)
//   icp_grid_cell (0 = auto)  icp_max_rmse (0.08)  icp_min_inlier_ratio (0.3)  ransac_seed (12345)
//   random_seed (-1 = time(0) like the reference)  use_voxel_grid (no)  gpu_device (0)
#include "ParameterReader.h"
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <set>

using namespace std;

ParameterReader *g_pParaReader = 0;
double camera_fx = 525.0, camera_fy = 525.0, camera_cx = 319.5, camera_cy = 235.5, camera_factor = 1000.0;

static string trim(const string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == string::npos ? string() : s.substr(a, b - a + 1);
}

ParameterReader::ParameterReader(const string &para_file) : _ok(false)
{
    cout << "init parameterReader, file addr = " << para_file << endl;
    ifstream fin(para_file.c_str());
    if (!fin) { cerr << "cannot open " << para_file << endl; return; }
    string line;
    while (getline(fin, line)) {
        size_t hash = line.find('#');
        if (hash != string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty() || line[0] == '%' || line == "---") continue;
        size_t colon = line.find(':');
        if (colon == string::npos) continue;
        string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (val.size() >= 2 && ((val[0] == '"' && val[val.size() - 1] == '"') || (val[0] == '\'' && val[val.size() - 1] == '\'')))
            val = val.substr(1, val.size() - 2);
        _kv[key] = val;
    }
    if (_kv.count("start_index") && _kv.count("end_index") && atoi(_kv["end_index"].c_str()) < atoi(_kv["start_index"].c_str())) {
        cerr << "end index should be larger than start index." << endl;   // reference :38-42
        return;
    }
    if (_kv.count("camera_fx")) camera_fx = atof(_kv["camera_fx"].c_str());
    if (_kv.count("camera_fy")) camera_fy = atof(_kv["camera_fy"].c_str());
    if (_kv.count("camera_cx")) camera_cx = atof(_kv["camera_cx"].c_str());
    if (_kv.count("camera_cy")) camera_cy = atof(_kv["camera_cy"].c_str());
    if (_kv.count("camera_factor")) camera_factor = atof(_kv["camera_factor"].c_str());
    _ok = true;
}

string ParameterReader::raw(const string &key, const string &def) const
{
    map<string, string>::const_iterator it = _kv.find(key);
    return it == _kv.end() ? def : it->second;
}

string ParameterReader::GetPara(const string &para_name)
{
    // typed like the reference's members: ints, doubles, strings
    static const char *ints[] = {"start_index", "end_index", "step_time", "optimize_step", "max_planes", "loopclosure_frames",
                                 "lost_frames", "loop_closure_inliers", 0};
    static const char *doubles[] = {"match_min_dist", "max_pos_change", "error_threshold", "grid_leaf", "distance_threshold",
                                    "plane_percent", "min_error_plane", "loop_closure_error", "error_odometry", "ransac_accuracy",
                                    "z_filter", 0};
    static const char *strings[] = {"data_source", "detector_name", "descriptor_name", "robust_kernel", "loop_closure_detection",
                                    "use_odometry", 0};
    for (int i = 0; ints[i]; ++i) if (para_name == ints[i]) return num2string(atoi(raw(para_name, "0").c_str()));
    for (int i = 0; doubles[i]; ++i) if (para_name == doubles[i]) return num2string(atof(raw(para_name, "0").c_str()));
    for (int i = 0; strings[i]; ++i) if (para_name == strings[i]) return raw(para_name, "");
    // optional keys of the ICP backend, with defaults
    struct Opt { const char *name; const char *def; };
    static const Opt opts[] = {{"icp_iterations", "10"}, {"icp_max_corr_dist", "0.2"}, {"icp_estimator", "plane"}, {"icp_search", "grid"},
                               {"icp_grid_cell", "0"}, {"icp_max_rmse", "0.08"}, {"icp_min_inlier_ratio", "0.3"}, {"ransac_seed", "12345"},
                               {"random_seed", "-1"}, {"use_voxel_grid", "no"}, {"gpu_device", "0"}, {0, 0}};
    for (int i = 0; opts[i].name; ++i) if (para_name == opts[i].name) return raw(para_name, opts[i].def);
    cerr << "Unknown parameter: " << para_name << endl;
    return string("unknown_para_name");
}
