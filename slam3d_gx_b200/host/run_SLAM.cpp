// run_SLAM.cpp -- the reference's driver (src/run_SLAM.cpp:11-44) on top of the B200 registration backend:
//   run_SLAM [loops]   reads ./parameters.yaml, registers frame after frame, writes ./data/final.g2o,
//                      ./data/final_after.g2o, ./data/keyframe.txt (+ lc.txt, lost.txt, error_of_transform.log)
#include "GraphicEnd.h"
#include <cstdlib>
#include <iostream>
using namespace std;

void usage() { cout << "usage: run_SLAM loops" << endl; }

int main(int argc, char **argv)
{
    int nloops = 3;                         // the reference leaves this uninitialised when argc > 2 (run_SLAM.cpp:13-28)
    if (argc < 2) usage();
    GraphicEnd *pGraphicEnd = new GraphicEnd();
    SLAMEnd *pSLAMEnd = new SLAMEnd();
    pGraphicEnd->init(pSLAMEnd);
    pSLAMEnd->init(pGraphicEnd);
    if (argc >= 2) nloops = atoi(argv[1]);
    for (int i = 0; i < nloops; i++) pGraphicEnd->run();
    cout << "Total KeyFrame: " << pGraphicEnd->_keyframes.size() << endl;
    pSLAMEnd->globalOptimizer.save("./data/final.g2o");
    pGraphicEnd->saveFinalResult("./data/final.pcd");
    delete pGraphicEnd;
    delete pSLAMEnd;
    return 0;
}
