// const.h -- constants shared by the host shell (counterpart of reference src/const.h:15,20,96-112;
// only the three items the front end actually uses are kept: the parameter file path, the camera
// intrinsics globals and the terminal colour macros).
#pragma once
#include <string>

const std::string parameter_file_addr = "./parameters.yaml";

// camera intrinsics, filled by ParameterReader (reference src/ParameterReader.cpp:9)
extern double camera_fx, camera_fy, camera_cx, camera_cy, camera_factor;

#define RESET "\033[0m"
#define RED "\033[31m"
#define GREEN "\033[32m"
#define YELLOW "\033[33m"
#define BOLDRED "\033[1m\033[31m"
#define BOLDGREEN "\033[1m\033[32m"
#define BOLDYELLOW "\033[1m\033[33m"
