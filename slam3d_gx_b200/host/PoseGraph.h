// PoseGraph.h -- stand-in for the g2o::SparseOptimizer held by SLAMEnd (reference src/GraphicEnd.h:223-256).
// The pose-graph optimiser is not on the hot path (out of scope) and g2o is not
// installable here; what IS needed is the on-disk result of the path: the g2o text file
// (VERTEX_SE3:QUAT / FIX / EDGE_SE3:QUAT with the 21 upper-triangular information entries) that the
// reference writes with globalOptimizer.save() (src/run_SLAM.cpp:36, src/GraphicEnd.cpp:680) and that
// g2o_viewer / generateTrajectory / saveOutput consume.  The class keeps the method names the front end calls.
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
    // No optimiser behind it: returns 0 iterations done (see header comment).
    int optimize(int) { return 0; }
    bool save(const char *filename) const;
    const std::vector<VertexSE3> &vertices() const { return _vertices; }
    const std::vector<EdgeSE3> &edges() const { return _edges; }

 private:
    std::vector<VertexSE3> _vertices;
    std::vector<EdgeSE3> _edges;
};
