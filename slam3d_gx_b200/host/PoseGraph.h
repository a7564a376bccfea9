// PoseGraph.h -- stand-in for the g2o::SparseOptimizer held by SLAMEnd (reference src/GraphicEnd.h:223-256).
// The pose-graph optimiser is not on the hot path (out of scope) and g2o is not
// installable here; what IS needed is the on-disk result of the path: the g2o text file
// (VERTEX_SE3:QUAT / FIX / EDGE_SE3:QUAT with the 21 upper-triangular information entries) that the
// reference writes with globalOptimizer.save() (src/run_SLAM.cpp:36, src/GraphicEnd.cpp:680) and reads back with
// globalOptimizer.load() in saveOutput / generateTrajectory (src/saveOutput.cpp:30, src/generateTrajectory.cpp:29),
// and that g2o_viewer consumes.  The class keeps the method names the front end and those tools call.
// optimize() does NOT run a least-squares solver: it only propagates vertex estimates along a spanning tree of the
// edges from the fixed vertex (what g2o does as computeInitialGuess before its first iteration), so that a graph
// whose new vertices were created at Identity (reference src/GraphicEnd.cpp:324) has usable poses on disk.
#pragma once
#include <string>
#include <vector>
#include "Pose.h"

struct VertexSE3 {
    int id; Isometry3d estimate; bool fixed;
    VertexSE3() : id(0), fixed(false) {}
    void setId(int i) { id = i; }
    void setEstimate(const Isometry3d &T) { estimate = T; }
    void setFixed(bool f) { fixed = f; }
    // x y z qx qy qz qw, the layout of g2o's VertexSE3::get/setEstimateData (reference src/generateTrajectory.cpp:61-65)
    void getEstimateData(double d[7]) const { d[0] = estimate(0, 3); d[1] = estimate(1, 3); d[2] = estimate(2, 3); estimate.quaternion(d + 3); }
    void setEstimateData(const double d[7]) { estimate = Isometry3d::fromQuaternion(d + 3, d); }
};

struct EdgeSE3 {
    int from, to; Isometry3d measurement; double information[6][6]; bool robust;
    EdgeSE3();
    void setVertices(int i, int j) { from = i; to = j; }
    void setMeasurement(const Isometry3d &T) { measurement = T; }
    void setInformationDiagonal(double v);
    void setRobustKernel(bool on) { robust = on; }   // Cauchy kernel flag (reference src/GraphicEnd.h:245); not serialised by g2o either
};

class SparseOptimizer
{
 public:
    bool addVertex(const VertexSE3 &v) { _vertices.push_back(v); return true; }
    bool addEdge(const EdgeSE3 &e) { _edges.push_back(e); return true; }
    const VertexSE3 *vertex(int id) const;
    void setVerbose(bool) {}
    bool initializeOptimization() { return true; }
    // No least-squares optimiser behind it (see header comment): spanning-tree propagation of the estimates from the fixed
    // vertices; returns the number of vertices that received an estimate this way.
    int optimize(int iterations);
    bool save(const char *filename) const;
    // g2o text reader (VERTEX_SE3:QUAT / FIX / EDGE_SE3:QUAT; other record types are skipped): replaces the graph
    bool load(const char *filename);
    void clear() { _vertices.clear(); _edges.clear(); }
    VertexSE3 *vertex(int id);
    const std::vector<VertexSE3> &vertices() const { return _vertices; }
    const std::vector<EdgeSE3> &edges() const { return _edges; }

 private:
    std::vector<VertexSE3> _vertices;
    std::vector<EdgeSE3> _edges;
};
