// PCD.h -- minimal reader for the clouds the reference loads with pcl::io::loadPCDFile (src/GraphicEnd.cpp:279-281):
// PCD v0.7, FIELDS x y z rgba, SIZE 4 4 4 4, DATA binary (16 bytes per point; header of reference
// data/exp1/pcd/1.pcd) or DATA ascii.  Rows come back as 4 floats (x,y,z,rgba bits) = the stride-4 layout
// s3d_cloud_upload copies straight to the device.
#pragma once
#include <string>
#include <vector>

// returns false on failure; xyzw gets 4 floats per point
bool loadPCDFile(const std::string &path, std::vector<float> &xyzw, int &n_points);
bool savePCDFileBinary(const std::string &path, const float *xyzw, int n_points);
