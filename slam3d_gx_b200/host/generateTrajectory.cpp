// generateTrajectory.cpp -- the reference's trajectory export (src/generateTrajectory.cpp:17-84):
//   generateTrajectory keyframe.txt final.g2o
// For every key frame (`id frame_index`, reference src/GraphicEnd.cpp:673-679) the time stamp of that frame -- first token
// of line `frame_index` of <data_source>/associate.txt, which is what the reference's getline/jump bookkeeping (:52-58,74)
// arrives at -- and the vertex estimate as `x y z qx qy qz qw` (VertexSE3::getEstimateData, :61-65) go to ./trajectory.txt
// in the TUM format `timestamp tx ty tz qx qy qz qw` (:66-71).  CPU only: nothing here touches the device.
#include "ParameterReader.h"
#include "PoseGraph.h"
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>
using namespace std;

static int usage() { cout << "generateTrajectory keyframe.txt final.g2o" << endl; return 0; }

int main(int argc, char **argv)
{
    if (argc != 3) { usage(); return -1; }
    g_pParaReader = new ParameterReader(parameter_file_addr);
    SparseOptimizer opt;
    if (!opt.load(argv[2])) { cout << "file does not exist" << endl; return -1; }
    ifstream fin(argv[1]);
    if (!fin) { cout << "file does not exist" << endl; return -1; }
    vector<string> stamps;
    {
        ifstream asso((g_pParaReader->GetPara("data_source") + string("/associate.txt")).c_str());
        string line;
        while (getline(asso, line)) { istringstream is(line); string t; is >> t; stamps.push_back(t); }
    }
    ofstream fout("trajectory.txt");
    int id, frame;
    while (fin >> id >> frame) {
        const VertexSE3 *pv = opt.vertex(id);
        if (pv == NULL) continue;                                      // reference :59-60
        double data[7];
        pv->getEstimateData(data);
        const string timestamp = (frame >= 1 && frame <= (int)stamps.size()) ? stamps[frame - 1] : string();
        fout << timestamp << " ";
        for (int i = 0; i < 7; i++) fout << data[i] << " ";            // reference :66-69 (default stream precision)
        fout << endl;
    }
    cout << "trajectory saved." << endl;
    delete g_pParaReader;
    return 0;
}
