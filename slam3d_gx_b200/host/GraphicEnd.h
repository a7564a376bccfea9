// GraphicEnd.h -- the front end of gaoxiang12/slam3d_gx with its public surface kept
// (GraphicEnd(), init(SLAMEnd*), run(), saveFinalResult(string), _keyframes; SLAMEnd::init, globalOptimizer.save --
// as used by reference src/run_SLAM.cpp:21-38), re-based on the B200 registration library:
//   extractPlanesAndGenerateImage  (reference src/GraphicEnd.cpp:353-430)  -> s3d_segment_planes
//   multiPnP                        (reference src/GraphicEnd.cpp:557-659)  -> s3d_register_pair  (ICP instead of SIFT + PnP-RANSAC)
//   loopClosure / lostRecovery      (reference :685-762 / :764-838)         -> one s3d_register_batch per sweep
// Frame loop, keyframe policy, gates and output files follow the reference line by line; what is not on the
// path (feature detection, GUI, image painting, g2o optimisation) is left out.
#pragma once
#include <atomic>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include "../../include/slam3d_b200.h"
#include "ParameterReader.h"
#include "Pose.h"
#include "PoseGraph.h"

class GraphicEnd;
class SLAMEnd;

// reference src/GraphicEnd.h:41-49 minus the feature members; every plane of a frame refers to the frame's cloud
struct PLANE {
    float coff[4];              // a,b,c,d with d >= 0
    int inliers;
    const s3d_cloud *cloud;     // device cloud of the frame this plane was extracted from (not owned)
    PLANE() : inliers(0), cloud(0) { coff[0] = coff[1] = coff[2] = coff[3] = 0.f; }
};

struct KEYFRAME {               // reference src/GraphicEnd.h:51-57
    int id;
    int frame_index;
    std::vector<PLANE> planes;
    std::vector<int> connect;
    KEYFRAME() : id(0), frame_index(0) {}
};

struct RESULT_OF_MULTIPNP {     // reference src/GraphicEnd.h:59-69
    RESULT_OF_MULTIPNP() : norm(0.0), inliers(0) {}
    Isometry3d T;
    double norm;
    int inliers;
};

class GraphicEnd
{
 public:
    GraphicEnd();
    virtual ~GraphicEnd();

    virtual void init(SLAMEnd *pSLAMEnd);
    virtual int run();
    virtual int readimage();
    virtual void generateKeyFrame(Isometry3d T);
    virtual void saveFinalResult(std::string fileaddr);

    // plane extraction on the device; the returned planes refer to `cloud`
    std::vector<PLANE> extractPlanesAndGenerateImage(s3d_cloud *cloud);
    // registration of frame 1 (plane1's cloud) onto frame 2 (plane2's cloud); T == Identity signals failure
    virtual RESULT_OF_MULTIPNP multiPnP(std::vector<PLANE> &plane1, std::vector<PLANE> &plane2, bool loopclosure = false,
                                        int frame_index = 0, int minimum_inliers = 12);
    // the same for many frame-1 candidates against one frame 2, in one device call
    std::vector<RESULT_OF_MULTIPNP> multiPnPBatch(const std::vector<std::vector<PLANE> *> &plane1, std::vector<PLANE> &plane2,
                                                  int minimum_inliers);

    void loopClosure();
    void displayLC(int frame1, int frame2, double norm, int inliers);
    void findMoreLoops();
    bool check(int frame1, int frame2);
    std::vector<int> checknearby(int source, int target);
    void lostRecovery();

 public:
    SLAMEnd *_pSLAMEnd;
    Isometry3d _robot, _kf_pos;
    std::vector<KEYFRAME> _keyframes;
    KEYFRAME _currKF, _present, _last;
    s3d_cloud *_currCloud;          // cloud of the frame being processed (owned by _clouds)

    int _lost;
    int _index;
    std::string _pclPath;
    double _distance_threshold, _percent, _max_pos_change, _error_threshold;
    int _max_planes;
    bool _loop_closure_detection;
    int _loopclosure_frames;
    double _loop_closure_error;
    int _lost_frames;
    double _z_filter;
    int _loop_closure_inliers;
    int _moreLoops;

    // backend
    s3d_ctx *_ctx;
    s3d_icp_params _icp;
    s3d_plane_params _seg;
    double _icp_max_rmse, _icp_min_inlier_ratio;
    // Device clouds that are still referenced: key frames keep theirs resident in HBM (they are loop-closure sources and
    // targets for the rest of the run), a frame that did not become a key frame is freed -- cloud, normals, labels and the
    // search index it got as a registration target -- as soon as _present/_last/_currKF have moved on (releaseUnused()).
    std::vector<s3d_cloud *> _clouds;
    size_t _peakClouds;                 // high-water mark of _clouds.size(): bounded by #key frames + 3
    // The NEXT frame: its PCD file is read and parsed by a worker thread while this frame is processed; as soon as the rows
    // are there the main thread (the only one that calls into the ctx) starts their upload on the copy stream
    // (s3d_cloud_upload_async), so disk read, PCIe copy and this frame's kernels overlap.
    s3d_cloud *_nextRaw;
    int _nextIndex;
    float *_pinned;                     // page-locked staging rows of the prefetched frame
    size_t _pinnedFloats;
    std::thread _reader;
    std::atomic<int> _readerState;      // 0 idle, 1 reading, 2 rows ready, 3 failed / no such file
    std::vector<float> _readRows;
    int _readN, _readIndex;
    std::stringstream ss;
    bool _have_guess, _use_guess;       // tracking: last successful key-frame -> frame pose as the next initial guess
    Isometry3d _guess;

 protected:
    RESULT_OF_MULTIPNP toResult(const s3d_result &r, int n_src, int minimum_inliers);
    void prefetch(int index);           // start the reader thread for frame `index`
    void issuePrefetched(bool wait);    // rows ready -> s3d_cloud_upload_async (wait: join the reader first)
    void releaseUnused();               // free every cloud no frame structure refers to any more
    void addEdge(int from, int to, const Isometry3d &T, double info, bool robust);
};

// reference src/GraphicEnd.h:223-256 without the g2o solver objects
class SLAMEnd
{
 public:
    SLAMEnd() : _pGraphicEnd(0) {}
    void init(GraphicEnd *p) { _pGraphicEnd = p; globalOptimizer.setVerbose(false); }

 public:
    GraphicEnd *_pGraphicEnd;
    SparseOptimizer globalOptimizer;
};
