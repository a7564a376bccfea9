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
    // test_host pcd <in.pcd> <out.bin>: dump what loadPCDFile reads (int32 n, then n rows of 4 floats) for the fixture tests
    if (argc == 4 && std::string(argv[1]) == "pcd") {
        std::vector<float> rows; int n = 0;
        if (!loadPCDFile(argv[2], rows, n)) { std::fprintf(stderr, "loadPCDFile failed\n"); return 2; }
        FILE *f = std::fopen(argv[3], "wb");
        REQUIRE(f);
        std::fwrite(&n, 4, 1, f); std::fwrite(rows.data(), 16, (size_t)n, f); std::fclose(f);
        return 0;
    }
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
    REQUIRE(Isometry3d().inverse().isIdentity());          // -0.0 in the inverse's translation is still the identity (reference :173 compares values)
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
    // --- g2o text reader: what save() wrote comes back (reference src/saveOutput.cpp:30, src/generateTrajectory.cpp:29 call opt.load)
    {
        Isometry3d T2 = T * T;
        VertexSE3 v2; v2.setId(2); opt.addVertex(v2);                      // created at Identity like reference GraphicEnd.cpp:324
        EdgeSE3 e2; e2.setVertices(1, 2); e2.setMeasurement(T); e2.setInformationDiagonal(100.0); e2.setRobustKernel(true); opt.addEdge(e2);
        EdgeSE3 e3; e3.setVertices(2, 0); e3.setMeasurement(T2.inverse()); e3.setInformationDiagonal(25.0); opt.addEdge(e3);
        std::string g2 = std::string(argv[2]) + "/t2.g2o";
        REQUIRE(opt.save(g2.c_str()));
        SparseOptimizer back;
        REQUIRE(back.load(g2.c_str()));
        REQUIRE(back.vertices().size() == 3 && back.edges().size() == 3);
        REQUIRE(back.vertex(0) && back.vertex(0)->fixed && !back.vertex(1)->fixed);
        for (int i = 0; i < 16; ++i) REQUIRE(std::fabs(back.vertex(1)->estimate.m[i] - T.m[i]) < 1e-8);    // %.9g on disk
        REQUIRE(back.edges()[1].from == 1 && back.edges()[1].to == 2 && back.edges()[2].information[3][3] == 25.0 && back.edges()[2].information[0][1] == 0.0);
        for (int i = 0; i < 16; ++i) REQUIRE(std::fabs(back.edges()[2].measurement.m[i] - T2.inverse().m[i]) < 1e-8);
        // estimate data round trip (x y z qx qy qz qw)
        double d[7]; back.vertex(1)->getEstimateData(d);
        VertexSE3 w; w.setEstimateData(d);
        for (int i = 0; i < 16; ++i) REQUIRE(std::fabs(w.estimate.m[i] - T.m[i]) < 1e-8);
        // spanning-tree propagation from the fixed vertex: vertex 2 (Identity on disk) becomes T * T
        REQUIRE(back.optimize(10) == 2);
        for (int i = 0; i < 16; ++i) REQUIRE(std::fabs(back.vertex(2)->estimate.m[i] - T2.m[i]) < 1e-7);
        REQUIRE(!back.load((std::string(argv[2]) + "/does_not_exist.g2o").c_str()));
    }
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
