// saveOutput.cpp -- the reference's map-fusion tool (src/saveOutput.cpp:13-99) on the B200 backend:
//   saveOutput keyframe.txt final.g2o [pass_z]
// reads ./parameters.yaml (grid_leaf, data_source), the key-frame list written by GraphicEnd::saveFinalResult
// (`id frame_index` per line, reference src/GraphicEnd.cpp:673-679) and the pose graph (SparseOptimizer::load, reference :30),
// loads every key frame's cloud, and runs the fusion loop of reference :47-95 -- voxel grid, z pass-through [0, pass_z],
// transform by the vertex estimate, append; voxel grid of the sum -- as ONE device call (s3d_map_fuse).  Writes result.pcd.
// (The reference's `while (!fin.eof())` processes the last key frame twice when the file ends with a newline; here every
// line is used once.)
#include "GraphicEnd.h"
#include "PCD.h"
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
using namespace std;

static void usage() { cout << "saveOutput keyframe.txt final.g2o [ pass_z ]" << endl; }

int main(int argc, char **argv)
{
    if (argc < 3) { usage(); return -1; }
    g_pParaReader = new ParameterReader(parameter_file_addr);
    SLAMEnd slam;
    slam.init(NULL);
    SparseOptimizer &opt = slam.globalOptimizer;
    if (!opt.load(argv[2])) { cerr << "cannot read " << argv[2] << endl; return -1; }
    ifstream fin(argv[1]);
    if (!fin) { cerr << "cannot read " << argv[1] << endl; return -1; }
    const double grid_leaf = atof(g_pParaReader->GetPara("grid_leaf").c_str());
    const string pclPath = g_pParaReader->GetPara("data_source") + "/pcd/";
    double z = 5.0;                                                    // reference :43-45
    if (argc == 4) z = atof(argv[3]);

    s3d_ctx *ctx = 0;
    if (s3d_create(&ctx, atoi(g_pParaReader->GetPara("gpu_device").c_str())) != S3D_OK) { cerr << "no CUDA device (no CPU fallback)" << endl; return -1; }
    vector<s3d_cloud *> clouds;
    vector<double> poses;
    int id, frame;
    while (fin >> id >> frame) {
        stringstream ss;
        ss << pclPath << frame << ".pcd";
        cout << "loading " << ss.str() << endl;
        const VertexSE3 *pv = opt.vertex(id);
        if (pv == NULL) { cout << "cannot find vertex: " << id << endl; continue; }      // reference :62-67
        vector<float> pts; int n = 0;
        if (!loadPCDFile(ss.str(), pts, n)) { cerr << "cannot read " << ss.str() << endl; return -1; }
        s3d_cloud *c = 0;
        if (s3d_cloud_upload(ctx, pts.data(), 4, n, &c) != S3D_OK) { cerr << s3d_last_error(ctx) << endl; return -1; }
        clouds.push_back(c);
        poses.insert(poses.end(), pv->estimate.m, pv->estimate.m + 16);
    }
    if (clouds.empty()) { cerr << "no key frame" << endl; return -1; }
    s3d_cloud *fused = 0;
    if (s3d_map_fuse(ctx, clouds.data(), poses.data(), (int)clouds.size(), (float)grid_leaf, (float)z, &fused) != S3D_OK) {
        cerr << s3d_last_error(ctx) << endl; return -1;
    }
    const int m = s3d_cloud_size(fused);
    vector<float> xyz((size_t)m * 3), rows((size_t)m * 4, 0.f);
    if (s3d_cloud_download(ctx, fused, xyz.data(), 0, 0) != S3D_OK) { cerr << s3d_last_error(ctx) << endl; return -1; }
    for (int i = 0; i < m; ++i) for (int k = 0; k < 3; ++k) rows[(size_t)4 * i + k] = xyz[(size_t)3 * i + k];
    savePCDFileBinary("result.pcd", rows.data(), m);                   // reference :96
    cout << "final result saved." << endl;
    for (size_t i = 0; i < clouds.size(); ++i) s3d_cloud_free(ctx, clouds[i]);
    s3d_cloud_free(ctx, fused);
    s3d_destroy(ctx);
    delete g_pParaReader;
    return 0;
}
