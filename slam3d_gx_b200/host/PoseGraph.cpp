#include "PoseGraph.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <map>
#include <deque>

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

VertexSE3 *SparseOptimizer::vertex(int id)
{
    for (size_t i = 0; i < _vertices.size(); ++i) if (_vertices[i].id == id) return &_vertices[i];
    return 0;
}

// Reads what save() (and g2o's SparseOptimizer::save) writes.  Edge information: 21 upper-triangular entries, row by row.
bool SparseOptimizer::load(const char *filename)
{
    std::ifstream fin(filename);
    if (!fin) return false;
    clear();
    std::string line;
    while (std::getline(fin, line)) {
        std::istringstream is(line);
        std::string tag;
        if (!(is >> tag)) continue;
        if (tag == "VERTEX_SE3:QUAT") {
            VertexSE3 v; double d[7];
            if (!(is >> v.id)) return false;
            for (int k = 0; k < 7; ++k) if (!(is >> d[k])) return false;
            v.setEstimateData(d);
            _vertices.push_back(v);
        } else if (tag == "FIX") {
            int id;
            while (is >> id) { VertexSE3 *v = vertex(id); if (v) v->fixed = true; }
        } else if (tag == "EDGE_SE3:QUAT") {
            EdgeSE3 e; double d[7];
            if (!(is >> e.from >> e.to)) return false;
            for (int k = 0; k < 7; ++k) if (!(is >> d[k])) return false;
            e.measurement = Isometry3d::fromQuaternion(d + 3, d);
            for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) {
                double v;
                if (!(is >> v)) return false;
                e.information[r][c] = e.information[c][r] = v;
            }
            _edges.push_back(e);
        }
    }
    return true;
}

// Spanning-tree propagation from the fixed vertices (breadth first, edges in insertion order): a vertex reached through
// edge (from -> to) gets estimate[to] = estimate[from] * measurement, through the reverse direction
// estimate[from] = estimate[to] * measurement^-1.  Odometry edges come first in insertion order at every vertex, so the
// key-frame chain is followed before loop-closure edges.  Vertices not connected to a fixed vertex keep their estimate.
int SparseOptimizer::optimize(int)
{
    std::map<int, std::vector<size_t> > adj;
    for (size_t i = 0; i < _edges.size(); ++i) { adj[_edges[i].from].push_back(i); adj[_edges[i].to].push_back(i); }
    std::map<int, bool> done;
    std::deque<int> queue;
    for (size_t i = 0; i < _vertices.size(); ++i) if (_vertices[i].fixed) { done[_vertices[i].id] = true; queue.push_back(_vertices[i].id); }
    int set = 0;
    while (!queue.empty()) {
        const int id = queue.front(); queue.pop_front();
        const VertexSE3 *v = vertex(id);
        if (!v) continue;
        const std::vector<size_t> &es = adj[id];
        for (size_t k = 0; k < es.size(); ++k) {
            const EdgeSE3 &e = _edges[es[k]];
            const int other = e.from == id ? e.to : e.from;
            if (done[other]) continue;
            VertexSE3 *o = vertex(other);
            if (!o) continue;
            o->estimate = e.from == id ? v->estimate * e.measurement : v->estimate * e.measurement.inverse();
            done[other] = true; ++set;
            queue.push_back(other);
        }
    }
    return set;
}
