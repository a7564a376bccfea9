#include "PoseGraph.h"
#include <cstdio>
#include <cstring>

EdgeSE3::EdgeSE3() : from(0), to(0), robust(false) { std::memset(information, 0, sizeof(information)); setInformationDiagonal(1.0); }

void EdgeSE3::setInformationDiagonal(double v)
{
    std::memset(information, 0, sizeof(information));
    for (int i = 0; i < 6; ++i) information[i][i] = v;
}

const VertexSE3 *SparseOptimizer::vertex(int id) const
{
    for (size_t i = 0; i < _vertices.size(); ++i) if (_vertices[i].id == id) return &_vertices[i];
    return 0;
}

// g2o text format: "VERTEX_SE3:QUAT id x y z qx qy qz qw", "FIX id", "EDGE_SE3:QUAT i j x y z qx qy qz qw" + 21 upper-triangular
// information entries, row by row.
bool SparseOptimizer::save(const char *filename) const
{
    FILE *f = std::fopen(filename, "w");
    if (!f) return false;
    for (size_t i = 0; i < _vertices.size(); ++i) {
        const VertexSE3 &v = _vertices[i];
        double q[4]; v.estimate.quaternion(q);
        std::fprintf(f, "VERTEX_SE3:QUAT %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", v.id, v.estimate(0, 3), v.estimate(1, 3), v.estimate(2, 3),
                     q[0], q[1], q[2], q[3]);
        if (v.fixed) std::fprintf(f, "FIX %d\n", v.id);
    }
    for (size_t i = 0; i < _edges.size(); ++i) {
        const EdgeSE3 &e = _edges[i];
        double q[4]; e.measurement.quaternion(q);
        std::fprintf(f, "EDGE_SE3:QUAT %d %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g", e.from, e.to, e.measurement(0, 3), e.measurement(1, 3),
                     e.measurement(2, 3), q[0], q[1], q[2], q[3]);
        for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) std::fprintf(f, " %.9g", e.information[r][c]);
        std::fprintf(f, "\n");
    }
    std::fclose(f);
    return true;
}
