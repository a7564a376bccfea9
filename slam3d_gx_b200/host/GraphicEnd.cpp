// GraphicEnd.cpp -- see GraphicEnd.h.  Control flow and constants mirror reference src/GraphicEnd.cpp; line
// references below point at the statement being mirrored.
#include "GraphicEnd.h"
#include "PCD.h"
#include <cstring>
#include <fstream>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iostream>

using namespace std;

#define S3D_CHECK(call) do { int rc__ = (call); if (rc__ != S3D_OK) { cerr << BOLDRED << "slam3d_b200: " #call " failed (" << rc__ << "): " \
    << s3d_last_error(_ctx) << RESET << endl; exit(1); } } while (0)

GraphicEnd::GraphicEnd() : _pSLAMEnd(0), _currCloud(0), _lost(0), _index(0), _moreLoops(0), _ctx(0), _peakClouds(0), _nextRaw(0), _nextIndex(-1), _pinned(0),
                           _pinnedFloats(0), _readerState(0), _readN(0), _readIndex(-1), _have_guess(false), _use_guess(false)
{
    g_pParaReader = new ParameterReader(parameter_file_addr);                       // :62
    int seed = atoi(g_pParaReader->GetPara("random_seed").c_str());
    srand(seed < 0 ? (unsigned int)time(0) : (unsigned int)seed);                   // :69 (seedable for reproducible runs)
    int dev = atoi(g_pParaReader->GetPara("gpu_device").c_str());
    if (s3d_create(&_ctx, dev) != S3D_OK) {
        cerr << BOLDRED << "cannot create the CUDA registration context on device " << dev << " (no CPU fallback)" << RESET << endl;
        exit(1);
    }
}

GraphicEnd::~GraphicEnd()
{
    if (_reader.joinable()) _reader.join();
    for (size_t i = 0; i < _clouds.size(); ++i) s3d_cloud_free(_ctx, _clouds[i]);
    if (_nextRaw) s3d_cloud_free(_ctx, _nextRaw);
    s3d_host_free(_ctx, _pinned);
    s3d_destroy(_ctx);
    delete g_pParaReader;
    g_pParaReader = 0;
}

void GraphicEnd::init(SLAMEnd *pSLAMEnd)
{
    cout << "Graphic end init..." << endl;
    _pSLAMEnd = pSLAMEnd;
    _index = atoi(g_pParaReader->GetPara("start_index").c_str());                                   // :82
    _pclPath = g_pParaReader->GetPara("data_source") + string("/pcd/");                            // :85
    _distance_threshold = atof(g_pParaReader->GetPara("distance_threshold").c_str());
    _error_threshold = atof(g_pParaReader->GetPara("error_threshold").c_str());
    _percent = atof(g_pParaReader->GetPara("plane_percent").c_str());
    _max_pos_change = atof(g_pParaReader->GetPara("max_pos_change").c_str());
    _max_planes = atoi(g_pParaReader->GetPara("max_planes").c_str());
    _loopclosure_frames = atoi(g_pParaReader->GetPara("loopclosure_frames").c_str());
    _loop_closure_detection = g_pParaReader->GetPara("loop_closure_detection") == string("yes");
    _loop_closure_error = atof(g_pParaReader->GetPara("loop_closure_error").c_str());
    _loop_closure_inliers = atoi(g_pParaReader->GetPara("loop_closure_inliers").c_str());
    _lost_frames = atoi(g_pParaReader->GetPara("lost_frames").c_str());
    _robot = _kf_pos = Isometry3d::Identity();
    _z_filter = atof(g_pParaReader->GetPara("z_filter").c_str());
    _lost = 0;                                                                                      // uninitialised in the reference (GraphicEnd.h:187)
    if (g_pParaReader->GetPara("use_odometry") == string("yes"))
        cerr << "use_odometry: the odometry path (reference :105-120) is outside the registration path and not provided" << endl;

    s3d_icp_params_default(&_icp);
    _icp.max_iterations = atoi(g_pParaReader->GetPara("icp_iterations").c_str());
    _icp.max_corr_dist = (float)atof(g_pParaReader->GetPara("icp_max_corr_dist").c_str());
    _icp.estimator = g_pParaReader->GetPara("icp_estimator") == string("svd") ? S3D_ESTIMATOR_SVD : S3D_ESTIMATOR_POINT_TO_PLANE;
    _icp.search = g_pParaReader->GetPara("icp_search") == string("brute") ? S3D_SEARCH_BRUTE : S3D_SEARCH_GRID;
    _icp.grid_cell = (float)atof(g_pParaReader->GetPara("icp_grid_cell").c_str());
    _icp_max_rmse = atof(g_pParaReader->GetPara("icp_max_rmse").c_str());
    _icp_min_inlier_ratio = atof(g_pParaReader->GetPara("icp_min_inlier_ratio").c_str());
    s3d_plane_params_default(&_seg);
    _seg.distance_threshold = (float)_distance_threshold;                                           // :364
    _seg.plane_percent = (float)_percent;
    _seg.max_planes = _max_planes;
    _seg.seed = (uint64_t)atoll(g_pParaReader->GetPara("ransac_seed").c_str());

    // first frame: it becomes key frame 0 and the fixed vertex of the graph (:122-146)
    readimage();
    _currKF.id = 0;
    _currKF.frame_index = _index;
    _currKF.planes = extractPlanesAndGenerateImage(_currCloud);
    _keyframes.push_back(_currKF);
    VertexSE3 v;
    v.setId(_currKF.id);
    v.setEstimate(_robot);
    v.setFixed(true);
    _pSLAMEnd->globalOptimizer.addVertex(v);
    _index++;
    cout << "********************" << endl;
}

int GraphicEnd::run()
{
    static ofstream errorfile("./data/error_of_transform.log");                                     // :153
    cout << "********************" << endl;
    _present.planes.clear();
    readimage();
    issuePrefetched(false);
    _present.planes = extractPlanesAndGenerateImage(_currCloud);                                    // :158
    issuePrefetched(false);

    // the reference registers from scratch every frame (features); an ICP tracks, so the previous frame's result
    // (same key frame) is the initial guess
    _use_guess = _have_guess;
    RESULT_OF_MULTIPNP result = multiPnP(_currKF.planes, _present.planes);                          // :168
    _use_guess = false;
    _have_guess = !result.T.isIdentity();
    if (_have_guess) _guess = result.T;
    Isometry3d T = result.T.inverse();                                                              // :170

    if (T.isIdentity()) {                                                                           // :173 lost
        errorfile << "9999" << endl;
        cout << BOLDRED "This frame lost" << RESET << endl;
        cout << "Matching last and present." << endl;
        RESULT_OF_MULTIPNP r = multiPnP(_last.planes, _present.planes);                             // :187
        if (r.T.isIdentity() || r.inliers < _loop_closure_inliers || r.norm > _loop_closure_error)
            _lost++;                                                                                // :188-189
        else {
            cout << BOLDGREEN << "add last as a new keyframe." << RESET << endl;
            _lost = 0;
            RESULT_OF_MULTIPNP rr = multiPnP(_currKF.planes, _last.planes);                         // :195
            _currKF.id++;
            _currKF.planes = _last.planes;
            _currKF.frame_index = _index - 1;
            _keyframes.push_back(_currKF);
            VertexSE3 v;
            v.setId(_currKF.id);
            v.setEstimate(Isometry3d::Identity());
            _pSLAMEnd->globalOptimizer.addVertex(v);
            addEdge(_currKF.id - 1, _currKF.id, rr.T.inverse(), 100.0, false);                      // :212-221
            generateKeyFrame(r.T.inverse());                                                        // :224
            _last = _present;
        }
    } else if (result.norm > _max_pos_change) {                                                     // :230 new key frame
        errorfile << result.norm << endl;
        _robot = T * _kf_pos;
        generateKeyFrame(T);
        _have_guess = false;                                                                        // new key frame: start from identity
        if (_loop_closure_detection) loopClosure();
        _lost = 0;
        _last = _present;
    } else {                                                                                        // :241 small motion
        errorfile << result.norm << endl;
        _robot = T * _kf_pos;
        _lost = 0;
        _last = _present;
    }
    if (_lost > _lost_frames) {                                                                     // :249
        cerr << "the robot lost. Perform lost recovery." << endl;
        lostRecovery();
        _last = _present;
    }
    _index++;
    _present.frame_index = _index;
    releaseUnused();
    return 1;
}

int GraphicEnd::readimage()
{
    cout << "loading image " << _index << endl;
    ss.str(""); ss.clear();
    ss << _pclPath << _index << ".pcd";                                                             // :279
    s3d_cloud *raw = 0, *c = 0;
    if (_readIndex == _index) issuePrefetched(true);       // the reader thread was given this frame: take its rows
    if (_nextRaw && _nextIndex == _index) {
        raw = _nextRaw;                  // uploaded while the previous frame was being registered
        _nextRaw = 0;
    } else {
        if (_reader.joinable()) _reader.join();
        _readerState = 0; _readIndex = -1;
        if (_nextRaw) { s3d_cloud_free(_ctx, _nextRaw); _nextRaw = 0; }      // the run jumped: the prefetched frame is not the one wanted
        vector<float> pts;
        int n = 0;
        if (!loadPCDFile(ss.str(), pts, n)) { cerr << "cannot read " << ss.str() << endl; exit(1); }
        S3D_CHECK(s3d_cloud_upload(_ctx, pts.data(), 4, n, &raw));
    }
    // PassThrough z in [0, z_filter] (:283-285) and, on request, VoxelGrid(grid_leaf) (:287-295), both on the device.
    // The reference always voxel-filters (its features do not need density); the ICP backend registers at full
    // density unless parameters.yaml says `use_voxel_grid: yes`.
    S3D_CHECK(s3d_cloud_passthrough_z(_ctx, raw, 0.0f, (float)_z_filter, &c));
    s3d_cloud_free(_ctx, raw);
    if (g_pParaReader->GetPara("use_voxel_grid") == string("yes")) {
        double grid = atof(g_pParaReader->GetPara("grid_leaf").c_str());
        s3d_cloud *v = 0;
        S3D_CHECK(s3d_cloud_voxel_grid(_ctx, c, (float)grid, &v));
        s3d_cloud_free(_ctx, c);
        c = v;
    }
    _clouds.push_back(c);
    if (_clouds.size() > _peakClouds) _peakClouds = _clouds.size();
    _currCloud = c;
    cout << "load ok." << endl;
    prefetch(_index + 1);
    return 0;
}

// Frame `index` is read and parsed on a worker thread while the frame just loaded is processed (the reference loads frame
// k+1 only when run() is called for it, src/GraphicEnd.cpp:266-281).  A missing file is not an error here: the caller
// decides where the sequence ends.  The worker touches neither the ctx nor the page-locked rows.
void GraphicEnd::prefetch(int index)
{
    if (_reader.joinable()) _reader.join();
    stringstream path;
    path << _pclPath << index << ".pcd";
    _readIndex = index;
    _readerState = 1;
    const string file = path.str();
    _reader = std::thread([this, file]() {
        int n = 0;
        const bool ok = loadPCDFile(file, _readRows, n) && n > 0;
        _readN = n;
        _readerState.store(ok ? 2 : 3, std::memory_order_release);
    });
}

// Main thread, between the stages of run(): when the reader has delivered the rows, copy them into the page-locked staging
// buffer and start the upload on the ctx copy stream; it crosses PCIe while the SMs extract planes / register.
void GraphicEnd::issuePrefetched(bool wait)
{
    if (wait && _reader.joinable()) _reader.join();
    if (_readerState.load(std::memory_order_acquire) < 2) return;
    if (_reader.joinable()) _reader.join();
    const bool ok = _readerState == 2;
    _readerState = 0;
    const int index = _readIndex;
    _readIndex = -1;
    if (!ok) return;
    if (_readRows.size() > _pinnedFloats) {
        // the staging rows of the frame before were consumed by readimage() (its pass-through returned with the stream drained)
        s3d_host_free(_ctx, _pinned);
        _pinned = 0; _pinnedFloats = 0;
        void *h = 0;
        if (s3d_host_alloc(_ctx, _readRows.size() * sizeof(float), &h) != S3D_OK) return;
        _pinned = (float *)h; _pinnedFloats = _readRows.size();
    }
    memcpy(_pinned, _readRows.data(), _readRows.size() * sizeof(float));
    if (_nextRaw) { s3d_cloud_free(_ctx, _nextRaw); _nextRaw = 0; }
    if (s3d_cloud_upload_async(_ctx, _pinned, 4, _readN, &_nextRaw) != S3D_OK) { _nextRaw = 0; return; }
    _nextIndex = index;
}

// Only key frames (and the three frame structures in flight) keep their clouds: everything else goes back to the device
// pool.  Without this every frame of a long run stayed resident with its search index (~56 MB each).
void GraphicEnd::releaseUnused()
{
    vector<const s3d_cloud *> used;
    used.push_back(_currCloud);
    const vector<PLANE> *live[3] = {&_currKF.planes, &_present.planes, &_last.planes};
    for (int k = 0; k < 3; ++k) for (size_t i = 0; i < live[k]->size(); ++i) used.push_back((*live[k])[i].cloud);
    for (size_t k = 0; k < _keyframes.size(); ++k) for (size_t i = 0; i < _keyframes[k].planes.size(); ++i) used.push_back(_keyframes[k].planes[i].cloud);
    size_t w = 0;
    for (size_t i = 0; i < _clouds.size(); ++i) {
        if (find(used.begin(), used.end(), (const s3d_cloud *)_clouds[i]) != used.end()) _clouds[w++] = _clouds[i];
        else s3d_cloud_free(_ctx, _clouds[i]);
    }
    _clouds.resize(w);
}

void GraphicEnd::addEdge(int from, int to, const Isometry3d &T, double info, bool robust)
{
    EdgeSE3 e;
    e.setVertices(from, to);
    e.setInformationDiagonal(info);          // 100 * I6 (:330-334)
    e.setMeasurement(T);
    e.setRobustKernel(robust);
    _pSLAMEnd->globalOptimizer.addEdge(e);
}

void GraphicEnd::generateKeyFrame(Isometry3d T)
{
    cout << BOLDGREEN << "GraphicEnd::generateKeyFrame" << RESET << endl;
    // the outgoing key frame stops being a registration target: only its points, normals and labels stay resident
    // (11 MB at 640x480); its search index (rebuilt on demand by check()) goes back to the pool
    if (!_currKF.planes.empty() && _currKF.planes[0].cloud) s3d_cloud_drop_index(_ctx, const_cast<s3d_cloud *>(_currKF.planes[0].cloud));
    _currKF.id++;                                                                                   // :308
    _currKF.planes = _present.planes;
    _currKF.frame_index = _index;
    _kf_pos = _robot;
    _keyframes.push_back(_currKF);
    VertexSE3 v;
    v.setId(_currKF.id);
    v.setEstimate(Isometry3d::Identity());                                                          // :324
    _pSLAMEnd->globalOptimizer.addVertex(v);
    addEdge(_currKF.id - 1, _currKF.id, T, 100.0, false);                                           // :327-337
}

vector<PLANE> GraphicEnd::extractPlanesAndGenerateImage(s3d_cloud *cloud)
{
    cout << "extracting planes" << endl;
    vector<PLANE> planes;
    s3d_plane out[S3D_MAX_PLANES];
    int n = 0;
    S3D_CHECK(s3d_segment_planes(_ctx, cloud, &_seg, out, &n));
    for (int i = 0; i < n; ++i) {
        PLANE p;
        for (int k = 0; k < 4; ++k) p.coff[k] = out[i].coef[k];
        p.inliers = out[i].inliers;
        p.cloud = cloud;
        cout << "Coff: " << p.coff[0] << "," << p.coff[1] << "," << p.coff[2] << "," << p.coff[3] << endl;   // :389
        planes.push_back(p);
    }
    cout << "Total planes: " << n << endl;
    return planes;
}

// s3d_result -> RESULT_OF_MULTIPNP with the reference's gates (:599-600, :621-624): on any failure T stays Identity
RESULT_OF_MULTIPNP GraphicEnd::toResult(const s3d_result &r, int n_src, int minimum_inliers)
{
    RESULT_OF_MULTIPNP result;
    result.inliers = r.inliers;
    if (r.status != S3D_PAIR_OK) return result;                                  // "object is empty" / degenerate
    if (r.inliers < minimum_inliers) return result;                              // :599-600
    if (r.inliers < _icp_min_inlier_ratio * n_src) return result;                // overlap gate of the ICP backend
    if (std::sqrt(r.fitness) > _icp_max_rmse) return result;                     // residual gate of the ICP backend
    cout << RED << "norm of Transform = " << r.norm << RESET << endl;
    result.norm = r.norm;                                                        // :618-620
    if (result.norm > _error_threshold) return result;                           // :621-624
    memcpy(result.T.m, r.T, sizeof(r.T));                                        // :645-655
    return result;
}

RESULT_OF_MULTIPNP GraphicEnd::multiPnP(vector<PLANE> &plane1, vector<PLANE> &plane2, bool, int, int minimum_inliers)
{
    cout << "solving multi PnP" << endl;
    if (plane1.empty() || plane2.empty() || !plane1[0].cloud || !plane2[0].cloud) {
        cout << "object is empty" << endl;                                       // :585-589
        return RESULT_OF_MULTIPNP();
    }
    s3d_result r;
    S3D_CHECK(s3d_register_pair(_ctx, plane1[0].cloud, plane2[0].cloud, _use_guess ? _guess.m : 0, &_icp, &r));
    cout << "Multipnp inliers = " << r.inliers << " (status " << r.status << ", iterations " << r.iterations << ", rmse " << std::sqrt(r.fitness) << ")" << endl;
    return toResult(r, s3d_cloud_size(plane1[0].cloud), minimum_inliers);
}

vector<RESULT_OF_MULTIPNP> GraphicEnd::multiPnPBatch(const vector<vector<PLANE> *> &plane1, vector<PLANE> &plane2, int minimum_inliers)
{
    vector<RESULT_OF_MULTIPNP> out(plane1.size());
    if (plane2.empty() || !plane2[0].cloud) return out;
    vector<const s3d_cloud *> src, tgt;
    vector<size_t> which;
    for (size_t i = 0; i < plane1.size(); ++i) {
        if (plane1[i]->empty() || !(*plane1[i])[0].cloud) continue;
        src.push_back((*plane1[i])[0].cloud); tgt.push_back(plane2[0].cloud); which.push_back(i);
    }
    if (src.empty()) return out;
    vector<s3d_result> res(src.size());
    S3D_CHECK(s3d_register_batch(_ctx, src.data(), tgt.data(), 0, (int)src.size(), &_icp, res.data()));
    for (size_t k = 0; k < which.size(); ++k) out[which[k]] = toResult(res[k], s3d_cloud_size(src[k]), minimum_inliers);
    return out;
}

void GraphicEnd::saveFinalResult(string)
{
    findMoreLoops();                                                                               // :664
    cout << "saving final result" << endl;
    SparseOptimizer &opt = _pSLAMEnd->globalOptimizer;
    opt.initializeOptimization();
    opt.optimize(atoi(g_pParaReader->GetPara("optimize_step").c_str()));                           // :669-670 (no optimiser: see PoseGraph.h)
    ofstream fout("./data/keyframe.txt");
    for (size_t i = 0; i < _keyframes.size(); i++) {
        cout << "keyframe " << i << " id = " << _keyframes[i].id << endl;
        fout << _keyframes[i].id << " " << _keyframes[i].frame_index << endl;                      // :673-679
    }
    opt.save("./data/final_after.g2o");                                                            // :680
    fout.close();
}

// Loop closure (:685-762): the candidates (the two key frames 3 and 4 back, then up to loopclosure_frames random older
// ones) are gathered first, registered against the current key frame in ONE batched device call (shared target), and the
// reference's accept/reject gates are applied in the original order.
void GraphicEnd::loopClosure()
{
    if (_keyframes.size() <= 3) return;                                                            // :687
    cout << "Checking loop closure." << endl;
    vector<int> cand; vector<bool> random_pick;
    for (int i = -3; i > -5; i--) {                                                                // :694
        int n = (int)_keyframes.size() + i;
        if (n >= 0) { cand.push_back(n); random_pick.push_back(false); } else break;
    }
    cout << "checking random frames" << endl;
    vector<int> checked;
    for (int i = 0; i < _loopclosure_frames; i++) {                                                // :729
        int frame = rand() % ((int)_keyframes.size() - 3);
        if (find(checked.begin(), checked.end(), frame) != checked.end()) continue;
        checked.push_back(frame);
        cand.push_back(frame); random_pick.push_back(true);
    }
    vector<vector<PLANE> *> p1;
    for (size_t k = 0; k < cand.size(); ++k) p1.push_back(&_keyframes[cand[k]].planes);
    vector<RESULT_OF_MULTIPNP> res = multiPnPBatch(p1, _currKF.planes, _loop_closure_inliers);
    for (size_t k = 0; k < cand.size(); ++k) {
        const RESULT_OF_MULTIPNP &result = res[k];
        if (result.T.isIdentity()) continue;                                                       // :703 / :739
        if (result.norm > _loop_closure_error) continue;
        if (result.inliers < _loop_closure_inliers) continue;
        Isometry3d T = result.T.inverse();
        if (random_pick[k]) displayLC(_keyframes[cand[k]].frame_index, _currKF.frame_index, result.norm, result.inliers);   // :746
        addEdge(_keyframes[cand[k]].id, _currKF.id, T, 100.0, true);                               // :711-722 / :748-759
        if (random_pick[k]) _keyframes.back().connect.push_back(cand[k]);                          // :761
    }
}

void GraphicEnd::lostRecovery()
{
    cout << BOLDYELLOW << "Lost Recovery..." << RESET << endl;                                     // :767
    _currKF.id++;
    _currKF.planes = _present.planes;
    _currKF.frame_index = _index;
    _kf_pos = _robot;
    ofstream fout("./data/lost.txt", ofstream::app);
    fout << _currKF.id << " " << _currKF.frame_index << endl;                                      // :775-777
    fout.close();
    _keyframes.push_back(_currKF);
    VertexSE3 v;
    v.setId(_currKF.id);
    v.setEstimate(Isometry3d::Identity());
    _pSLAMEnd->globalOptimizer.addVertex(v);
    // no edge to the previous key frame: the position is unknown (:791-792); brute-force sweep over ALL earlier key frames (:810-836)
    vector<vector<PLANE> *> p1;
    for (size_t i = 0; i + 1 < _keyframes.size(); i++) p1.push_back(&_keyframes[i].planes);
    vector<RESULT_OF_MULTIPNP> res = multiPnPBatch(p1, _currKF.planes, 12);
    for (size_t i = 0; i < res.size(); i++) {
        const RESULT_OF_MULTIPNP &result = res[i];
        if (result.T.isIdentity()) continue;
        if (result.inliers < _loop_closure_inliers) continue;
        if (result.norm > _loop_closure_error) continue;
        addEdge(_keyframes[i].id, _currKF.id, result.T.inverse(), 100.0, true);
        _keyframes.back().connect.push_back((int)i);
    }
    _lost = 0;
}

void GraphicEnd::displayLC(int frame1, int frame2, double norm, int inliers)
{
    static ofstream fout("./data/lc.txt");                                                         // :842
    fout << frame1 << " " << frame2 << " " << norm << " " << inliers << endl;                      // :861
}

void GraphicEnd::findMoreLoops()
{
    cout << "Find more loops" << endl;                                                             // :864
    _moreLoops = 0;
    for (size_t i = 0; i < _keyframes.size(); i++) {
        if (_keyframes[i].connect.size() == 0) continue;
        vector<int> checked;
        for (size_t j = 0; j < _keyframes[i].connect.size(); j++) {
            checked = checknearby((int)i, _keyframes[i].connect[j]);
            for (size_t k = 0; k < checked.size(); k++) checknearby(checked[k], (int)i);
        }
    }
    cout << BOLDRED << "Total " << _moreLoops << " loops found. " << RESET << endl;
    size_t live = 0, peak = 0;
    s3d_memory_stats(_ctx, &live, &peak, 0);
    cout << "Peak resident clouds: " << _peakClouds << " peak device bytes: " << peak << endl;
}

bool GraphicEnd::check(int frame1, int frame2)
{
    cout << YELLOW << "checking " << frame1 << ", " << frame2 << RESET << endl;                     // :889
    RESULT_OF_MULTIPNP result = multiPnP(_keyframes[frame1].planes, _keyframes[frame2].planes, true, _keyframes[frame1].frame_index,
                                         _loop_closure_inliers);
    // the search index this registration built on key frame `frame2` is not kept (only the current key frame stays a target)
    if (!_keyframes[frame2].planes.empty() && _keyframes[frame2].planes[0].cloud != (_currKF.planes.empty() ? 0 : _currKF.planes[0].cloud))
        s3d_cloud_drop_index(_ctx, const_cast<s3d_cloud *>(_keyframes[frame2].planes[0].cloud));
    if (result.T.isIdentity()) return false;
    if (result.norm > _loop_closure_error) return false;
    if (result.inliers < _loop_closure_inliers) return false;
    addEdge(_keyframes[frame1].id, _keyframes[frame2].id, result.T.inverse(), 100.0, true);
    _moreLoops++;
    return true;
}

vector<int> GraphicEnd::checknearby(int source, int target)
{
    cout << RED << "checking " << source << " and " << target << RESET << endl;                     // :919
    vector<int> checked;
    int index = target;
    while (index > 0) {
        index--;
        if (index == source) continue;
        if (check(source, index)) checked.push_back(index); else break;
    }
    index = target;
    while (index < (int)_keyframes.size() - 1) {
        index++;
        if (index == source) continue;
        if (check(source, index)) checked.push_back(index); else break;
    }
    return checked;
}
