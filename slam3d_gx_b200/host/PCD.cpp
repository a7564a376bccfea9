#include "PCD.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

bool loadPCDFile(const std::string &path, std::vector<float> &xyzw, int &n_points)
{
    std::ifstream fin(path.c_str(), std::ios::binary);
    if (!fin) return false;
    std::string line, data_kind;
    std::vector<std::string> fields;
    std::vector<int> sizes;
    int points = -1, width = 0, height = 1;
    while (std::getline(fin, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream is(line);
        std::string key; is >> key;
        if (key == "FIELDS") { std::string f; while (is >> f) fields.push_back(f); }
        else if (key == "SIZE") { int s; while (is >> s) sizes.push_back(s); }
        else if (key == "WIDTH") is >> width;
        else if (key == "HEIGHT") is >> height;
        else if (key == "POINTS") is >> points;
        else if (key == "DATA") { is >> data_kind; break; }
    }
    if (points < 0) points = width * height;
    if (fields.size() < 3 || fields[0] != "x" || fields[1] != "y" || fields[2] != "z") return false;
    if (sizes.empty()) sizes.assign(fields.size(), 4);
    size_t stride = 0;
    for (size_t i = 0; i < sizes.size(); ++i) stride += (size_t)sizes[i];
    n_points = points;
    xyzw.assign((size_t)points * 4, 0.f);
    if (data_kind == "binary") {
        if (sizes[0] != 4 || sizes[1] != 4 || sizes[2] != 4) return false;
        std::vector<char> buf(stride * (size_t)points);
        fin.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)fin.gcount() != buf.size()) return false;
        for (int i = 0; i < points; ++i) {
            std::memcpy(&xyzw[(size_t)i * 4], &buf[(size_t)i * stride], 12);
            if (stride >= 16) std::memcpy(&xyzw[(size_t)i * 4 + 3], &buf[(size_t)i * stride + 12], 4);
        }
        return true;
    }
    if (data_kind == "ascii") {
        for (int i = 0; i < points; ++i) {
            if (!std::getline(fin, line)) return false;
            std::istringstream is(line);
            is >> xyzw[(size_t)i * 4] >> xyzw[(size_t)i * 4 + 1] >> xyzw[(size_t)i * 4 + 2];
        }
        return true;
    }
    return false;   // binary_compressed is not produced by the reference's tools
}

bool savePCDFileBinary(const std::string &path, const float *xyzw, int n)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
                    "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n", n, n);
    std::fwrite(xyzw, 16, (size_t)n, f);
    std::fclose(f);
    return true;
}
