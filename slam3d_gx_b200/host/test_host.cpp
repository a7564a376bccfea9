// test_host.cpp -- CPU-only checks of the host shell pieces that do not touch the GPU (run by tests/test_host_shell.py).
#include "ParameterReader.h"
#include "PCD.h"
#include "PoseGraph.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char **argv)
{
    REQUIRE(argc == 3);
    // --- ParameterReader on a reference-style parameters.yaml
    ParameterReader pr(argv[1]);
    REQUIRE(pr.ok());
    REQUIRE(pr.GetPara("data_source") == "/tmp/some dataset");
    REQUIRE(pr.GetPara("detector_name") == "SIFT");
    REQUIRE(pr.GetPara("start_index") == "1");
    REQUIRE(pr.GetPara("end_index") == "2800");
    REQUIRE(pr.GetPara("distance_threshold") == "0.08");      // "0.080" in the file: numbers are re-formatted like the reference's num2string
    REQUIRE(pr.GetPara("plane_percent") == "0.2");
    REQUIRE(pr.GetPara("max_planes") == "3");
    REQUIRE(pr.GetPara("loop_closure_detection") == "yes");
    REQUIRE(pr.GetPara("loopclosure_frames") == "30");
    REQUIRE(pr.GetPara("z_filter") == "7");
    REQUIRE(pr.GetPara("optimize_step") == "200");
    REQUIRE(std::fabs(camera_fx - 517.0) < 1e-12 && std::fabs(camera_cy - 255.3) < 1e-12 && std::fabs(camera_factor - 5000.0) < 1e-12);
    REQUIRE(pr.GetPara("icp_iterations") == "10");            // optional key absent from the file: default
    REQUIRE(pr.GetPara("icp_estimator") == "plane");
    REQUIRE(pr.GetPara("no_such_key") == "unknown_para_name"); // reference src/ParameterReader.cpp:121-122
    // --- Pose algebra
    Isometry3d T;
    REQUIRE(T.isIdentity());
    double a = 0.3;
    T(0, 0) = std::cos(a); T(0, 1) = -std::sin(a); T(1, 0) = std::sin(a); T(1, 1) = std::cos(a); T(0, 3) = 1; T(1, 3) = -2; T(2, 3) = 0.5;
    Isometry3d I = T * T.inverse();
    for (int i = 0; i < 16; ++i) REQUIRE(std::fabs(I.m[i] - Isometry3d().m[i]) < 1e-14);
    REQUIRE(!T.isIdentity());
    double q[4]; T.quaternion(q);
    REQUIRE(std::fabs(q[2] - std::sin(a / 2)) < 1e-14 && std::fabs(q[3] - std::cos(a / 2)) < 1e-14 && std::fabs(q[0]) < 1e-15);
    // --- g2o text output
    SparseOptimizer opt;
    VertexSE3 v0; v0.setId(0); v0.setFixed(true); opt.addVertex(v0);
    VertexSE3 v1; v1.setId(1); v1.setEstimate(T); opt.addVertex(v1);
    EdgeSE3 e; e.setVertices(0, 1); e.setMeasurement(T); e.setInformationDiagonal(100.0); opt.addEdge(e);
    std::string g2o = std::string(argv[2]) + "/t.g2o";
    REQUIRE(opt.save(g2o.c_str()));
    REQUIRE(opt.vertex(1) && opt.vertex(1)->id == 1 && !opt.vertex(7));
    // --- PCD round trip
    std::vector<float> pts;
    for (int i = 0; i < 5; ++i) { pts.push_back(0.1f * i); pts.push_back(-1.f * i); pts.push_back(2.f + i); pts.push_back(0.f); }
    std::string pcd = std::string(argv[2]) + "/t.pcd";
    REQUIRE(savePCDFileBinary(pcd, pts.data(), 5));
    std::vector<float> back; int n = 0;
    REQUIRE(loadPCDFile(pcd, back, n) && n == 5);
    for (int i = 0; i < 5; ++i) for (int k = 0; k < 3; ++k) REQUIRE(back[4 * i + k] == pts[4 * i + k]);
    std::puts("host shell ok");
    return 0;
}
