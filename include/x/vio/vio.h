// x::VIO of the B200 back end for callers that deliver feature MATCHES (jpl-x/x_multi_agent include/x/vio/vio.h:43-246,
// src/x/vio/vio.cpp): the reference's public entry points over the operator classes of include/x/xb200_binding.hpp --
//   setUp (vio.cpp:113-215), initAtTime (:54-111), processImu (:343-370, incl. the self-initialisation from the first 51
//   accelerometer samples), processMatchesMeasurement (:274-323) with importMatches (:372-434),
//   setLastRangeMeasurement / setLastSunAngleMeasurement (:217-224), loadParamsFromYaml (:576-707),
//   computeSLAMCartesianFeaturesForState (:328-332).
// What runs where: parameters, match import and track management are host code (TrackManager: xb_tm_* of libxb200.so),
// everything from the track lists on is the device filter.  Out of scope ("out"): processImageMeasurement
// (Tracker / KLT on pixels), place recognition.  The Python mirror of this class is x_multi_agent_b200/vio.py.
#pragma once
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "../xb200_binding.hpp"

namespace x {
namespace fsm = std::filesystem;

class VIO {
 public:
  VIO() : ekf_{Ekf(vio_updater_)} {}   // vio.cpp:40
  VIO(const VIO&) = delete;            // ekf_ refers to this object's vio_updater_
  VIO& operator=(const VIO&) = delete;

  [[nodiscard]] bool isInitialized() const { return initialized_; }

  /** vio.cpp:113-215.  `max_tracks` (capacity of one update) and `device` (CUDA ordinal) are this back end's additions. */
  void setUp(const Params& params, int max_tracks = 1024, int device = 0) {
    params_ = params;
    self_init_start_ = params_.self_init_start_;
    initialize_start_ = self_init_start_;
    camera_ = Camera(params_.cam_fx, params_.cam_fy, params_.cam_cx, params_.cam_cy, params_.cam_s,
                     static_cast<unsigned int>(params_.img_width), static_cast<unsigned int>(params_.img_height));
    if (params_.min_track_length > params_.n_poses_max)
      throw std::invalid_argument("'min_track_length' cannot be larger than 'n_poses_max'");
    tracker_ = Tracker();
    // minimum MSCKF baseline in the normal plane (vio.cpp:163-168)
    msckf_baseline_x_n_ = params_.msckf_baseline / (params_.img_width * params_.cam_fx);
    msckf_baseline_y_n_ = params_.msckf_baseline / (params_.img_height * params_.cam_fy);
    track_manager_ = TrackManager(camera_, msckf_baseline_x_n_, msckf_baseline_y_n_);
    const int n_poses_state = params_.n_poses_max;
    const int n_features_state = params_.n_slam_features_max;
    state_manager_ = StateManager(n_poses_state, n_features_state);
    const Vector3 g = params_.g;
    ImuNoise imu_noise;
    imu_noise.n_w = params_.n_w;
    imu_noise.n_bw = params_.n_bw;
    imu_noise.n_a = params_.n_a;
    imu_noise.n_ba = params_.n_ba;
    double sigma_landmark = 0.0, ci_msckf_w = -1.0, ci_slam_w = -1.0;
#ifdef MULTI_UAV
    sigma_landmark = params_.sigma_landmark;
    ci_msckf_w = params_.ci_msckf_w;
    ci_slam_w = params_.ci_slam_w;
#endif
    vio_updater_ = VioUpdater(tracker_, state_manager_, track_manager_, params_.sigma_img, params_.sigma_range, params_.rho_0,
                              params_.sigma_rho_0, params_.min_track_length, sigma_landmark, ci_msckf_w, ci_slam_w,
                              params_.iekf_iter);
    const size_t state_buffer_sz = static_cast<size_t>(params_.state_buffer_size);
    const State default_state = State(n_poses_state, n_features_state);
    const double a_m_max = 50.0;
    const unsigned int delta_seq_imu = 1;
    const auto time_margin_bfr = 0.02;
    ekf_.set(vio_updater_, g, imu_noise, static_cast<int>(state_buffer_sz), default_state, a_m_max, delta_seq_imu,
             time_margin_bfr, max_tracks, device);
    // image point of the laser range finder (vio.cpp:288-294): constant for a camera
    Feature lrf_img_pt;
    lrf_img_pt.setXDist(static_cast<double>((camera_.getWidth() + 1) / 2.0));
    lrf_img_pt.setYDist(static_cast<double>((camera_.getHeight() + 1) / 2.0));
    lrf_img_pt_n_ = camera_.undistortAndNormalize(lrf_img_pt);
    initialized_ = false;
  }

  /** vio.cpp:54-111 */
  void initAtTime(const double& time) {
    initialized_ = false;
    initialize_start_ = self_init_start_;
    ekf_.lock();   // held to the end, as vio.cpp:55-109 does (the mutex of this binding is recursive)
    vio_updater_.track_manager_.clear();
    vio_updater_.state_manager_.clear();
    // initial IMU measurement: gravity reaction along the IMU +Z axis, no rotation
    const Vector3 a_m = -params_.g;
    const Vector3 w_m(0.0, 0.0, 0.0);
    // initial vision state estimates and uncertainties are all zero
    const int n_poses_state = params_.n_poses_max;
    const int n_features_state = params_.n_slam_features_max;
    const Matrix p_array = Matrix::Zero(n_poses_state * 3, 1);
    const Matrix q_array = Matrix::Zero(n_poses_state * 4, 1);
    const Matrix f_array = Matrix::Zero(n_features_state * 3, 1);
    const int n_err = kSizeCoreErr + n_poses_state * 6 + n_features_state * 3;
    Matrix cov = Matrix::Zero(n_err, n_err);
    const double deg = M_PI / 180.0;
    for (int i = 0; i < 3; ++i) {
      const double s[5] = {params_.sigma_dp(i), params_.sigma_dv(i), params_.sigma_dtheta(i) * deg, params_.sigma_dbw(i) * deg,
                           params_.sigma_dba(i)};
      for (int b = 0; b < 5; ++b) cov(3 * b + i, 3 * b + i) = s[b] * s[b];
    }
    const unsigned int dummy_seq = 0;
    State init_state(time, dummy_seq, params_.p, params_.v, params_.q, params_.b_w, params_.b_a, p_array, q_array, f_array, cov,
                     params_.q_ic, params_.p_ic, w_m, a_m);
    try {
      ekf_.initializeFromState(init_state);
    } catch (std::runtime_error& e) {
      std::cerr << "bad input: " << e.what() << std::endl;
    } catch (init_bfr_mismatch&) {
      std::cerr << "init_bfr_mismatch: the size of dynamic arrays in the initialization state match must match the size "
                   "allocated in the buffered states."
                << std::endl;
    }
    ekf_.unlock();
    initialized_ = true;
  }

  void setLastRangeMeasurement(const RangeMeasurement& range_measurement) { last_range_measurement_ = range_measurement; }
  void setLastSunAngleMeasurement(const SunAngleMeasurement& angle_measurement) { last_angle_measurement_ = angle_measurement; }

  /** vio.cpp:343-370 */
  std::optional<State> processImu(const double& timestamp, unsigned int seq, const Vector3& w_m, const Vector3& a_m) {
    if (initialize_start_) {
      if (imu_data_batch_.size() < 50) {
        imu_data_batch_.push_back(a_m);
        return std::nullopt;
      }
      imu_data_batch_.push_back(a_m);
      Vector3 avg_a(0.0, 0.0, 0.0);
      for (const auto& v : imu_data_batch_) avg_a += v;
      avg_a /= static_cast<double>(imu_data_batch_.size());
      const Vector3 g(0, 0, a_m.norm());
      params_.q = fromTwoVectors(avg_a, g);
      initAtTime(timestamp);
      imu_data_batch_.clear();
      initialize_start_ = false;
      return std::nullopt;
    }
    return ekf_.processImu(timestamp, seq, w_m, a_m);
  }

  /** vio.cpp:274-323.  The images of the reference carry pixels for its GUI; here they carry the tile grid. */
  std::optional<State> processMatchesMeasurement(const double& timestamp, unsigned int seq, const std::vector<double>& match_vector,
                                                 TiledImage& match_img, TiledImage& feature_img) {
    const auto timestamp_corrected = timestamp + params_.time_offset;
    // import matches (except for the first measurement: the previous image has to enter the sliding window first)
    MatchList matches;
    if (vio_updater_.state_manager_.poseSize()) matches = importMatches(match_vector, seq, match_img);
    last_range_measurement_.img_pt_n = lrf_img_pt_n_;
    VioMeasurement measurement(timestamp_corrected, seq, matches, feature_img, last_range_measurement_, last_angle_measurement_);
    vio_updater_.setMeasurement(measurement);
    auto updated_state = ekf_.processUpdateMeasurement();
    // the state carries the original image timestamp for identification in the output
    if (updated_state.has_value()) updated_state->setTime(timestamp);
    feature_img = vio_updater_.getFeatureImage();
    return updated_state;
  }
  /** The same call with the tile grid of the parameters (n_tiles_h x n_tiles_w). */
  std::optional<State> processMatchesMeasurement(const double& timestamp, unsigned int seq, const std::vector<double>& match_vector) {
    TiledImage match_img(static_cast<unsigned int>(params_.n_tiles_h), static_cast<unsigned int>(params_.n_tiles_w));
    TiledImage feature_img = match_img;
    return processMatchesMeasurement(timestamp, seq, match_vector, match_img, feature_img);
  }

  /** vio.cpp:328-332 */
  std::vector<Vector3> computeSLAMCartesianFeaturesForState(const State& state) {
    return vio_updater_.state_manager_.computeSLAMCartesianFeaturesForState(state);
  }

  /** vio.cpp:576-707 without cv::FileStorage: the flat `key: value` / `key: [a, b, ...]` files the reference ships
   *  (the `%YAML:1.0` directive and `---` lines OpenCV writes are skipped).  Vectors and quaternions (w, x, y, z) as there;
   *  a missing key keeps the default of Params. */
  Params loadParamsFromYaml(fsm::path& path) { return loadParamsFromYaml(static_cast<const fsm::path&>(path)); }
  /** The same for a temporary (the reference's README calls it with a string literal). */
  Params loadParamsFromYaml(const fsm::path& path) {
    std::ifstream in(path);
    if (!in) throw std::invalid_argument("cannot open parameter file " + path.string());
    std::map<std::string, std::vector<std::string>> doc;
    std::string line, pending_key, pending;
    auto flush_list = [&](const std::string& key, std::string body) {
      for (char& ch : body) if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
      std::istringstream ss(body);
      std::vector<std::string> items;
      for (std::string tok; ss >> tok;) items.push_back(tok);
      doc[key] = items;
    };
    while (std::getline(in, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      if (line.rfind("%YAML", 0) == 0 || line.rfind("---", 0) == 0) continue;
      if (!pending_key.empty()) {   // continuation of a flow sequence that spans lines
        pending += " " + line;
        if (line.find(']') != std::string::npos) { flush_list(pending_key, pending); pending_key.clear(); }
        continue;
      }
      const size_t colon = line.find(':');
      if (colon == std::string::npos) continue;
      std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
      if (key.empty()) continue;
      if (!val.empty() && val[0] == '[' && val.find(']') == std::string::npos) { pending_key = key; pending = val; continue; }
      if (!val.empty() && val[0] == '[') flush_list(key, val);
      else doc[key] = {unquote(val)};
    }
    Params params;
    auto num = [&](const char* key, auto& dst) {
      auto it = doc.find(key);
      if (it == doc.end() || it->second.empty() || it->second[0].empty()) return;
      const std::string& s = it->second[0];
      using T = std::decay_t<decltype(dst)>;
      if constexpr (std::is_same_v<T, bool>) dst = (s == "true" || s == "True" || s == "1");
      else dst = static_cast<T>(std::stod(s));
    };
    auto vec3 = [&](const char* key, Vector3& dst) {
      auto it = doc.find(key);
      if (it == doc.end()) return;
      if (it->second.size() != 3) throw std::invalid_argument(std::string("parameter '") + key + "' needs 3 values");
      dst = Vector3(std::stod(it->second[0]), std::stod(it->second[1]), std::stod(it->second[2]));
    };
    auto quat = [&](const char* key, Quaternion& dst) {
      auto it = doc.find(key);
      if (it == doc.end()) return;
      if (it->second.size() != 4) throw std::invalid_argument(std::string("parameter '") + key + "' needs 4 values");
      dst = Quaternion(std::stod(it->second[0]), std::stod(it->second[1]), std::stod(it->second[2]), std::stod(it->second[3]));
      dst.normalize();
    };
    vec3("p", params.p); vec3("v", params.v); quat("q", params.q); vec3("b_w", params.b_w); vec3("b_a", params.b_a);
    vec3("sigma_dp", params.sigma_dp); vec3("sigma_dv", params.sigma_dv); vec3("sigma_dtheta", params.sigma_dtheta);
    vec3("sigma_dbw", params.sigma_dbw); vec3("sigma_dba", params.sigma_dba);
    num("cam1_fx", params.cam_fx); num("cam1_fy", params.cam_fy); num("cam1_cx", params.cam_cx); num("cam1_cy", params.cam_cy);
    num("cam1_s", params.cam_s); num("cam1_img_height", params.img_height); num("cam1_img_width", params.img_width);
    vec3("cam1_p_ic", params.p_ic); quat("cam1_q_ic", params.q_ic);
    num("cam1_time_offset", params.time_offset); num("sigma_img", params.sigma_img); num("sigma_range", params.sigma_range);
    quat("q_sc", params.q_sc);
    vec3("w_s", params.w_s);
    if (doc.count("w_s")) params.w_s = params.w_s.normalized();
    num("n_a", params.n_a); num("n_ba", params.n_ba); num("n_w", params.n_w); num("n_bw", params.n_bw);
    if (auto it = doc.find("vocabulary_path"); it != doc.end() && !it->second.empty()) params.vocabulary_path = it->second[0];
    num("sigma_landmark", params.sigma_landmark); num("descriptor_scale_factor", params.descriptor_scale_factor);
    num("descriptor_pyramid", params.descriptor_pyramid); num("descriptor_patch_size", params.descriptor_patch_size);
    num("ci_msckf_w", params.ci_msckf_w); num("ci_slam_w", params.ci_slam_w); num("desc_type", params.desc_type);
    num("pr_score_thr", params.pr_score_thr); num("pr_desc_ratio_thr", params.pr_desc_ratio_thr);
    num("pr_desc_min_distance", params.pr_desc_min_distance);
    num("min_eig_thr", params.min_eig_thr); num("max_level", params.max_level); num("win_size_w", params.win_size_w);
    num("win_size_h", params.win_size_h); num("fast_detection_delta", params.fast_detection_delta);
    num("non_max_supp", params.non_max_supp); num("block_half_length", params.block_half_length); num("margin", params.margin);
    num("n_feat_min", params.n_feat_min); num("outlier_method", params.outlier_method);
    num("outlier_param1", params.outlier_param1); num("outlier_param2", params.outlier_param2);
    num("n_tiles_h", params.n_tiles_h); num("n_tiles_w", params.n_tiles_w); num("max_feat_per_tile", params.max_feat_per_tile);
    num("n_poses_max", params.n_poses_max); num("n_slam_features_max", params.n_slam_features_max); num("rho_0", params.rho_0);
    num("sigma_rho_0", params.sigma_rho_0); num("iekf_iter", params.iekf_iter); num("msckf_baseline", params.msckf_baseline);
    num("min_track_length", params.min_track_length); num("state_buffer_size", params.state_buffer_size);
    vec3("g", params.g);
    return params;
  }

#ifdef MULTI_UAV
  /** A correspondence between this agent's data and a peer's, as the reference's place recognition reports it
   *  (PlaceRecognition::findCorrespondences fills the tracker's match lists, vio.cpp:538-546): SLAM feature `current`
   *  of this agent is the peer's SLAM feature `received`; MSCKF track with id `current` is the peer's MSCKF track number
   *  `received` of the list it sent.  Descriptor matching on pixels is the front end's job (out of scope): the caller
   *  delivers the correspondences, as it delivers visual matches. */
  struct Correspondence { unsigned long long current; int received; };

  /** vio.cpp:440-452 */
  void getDataToSend(std::shared_ptr<SimpleState>& state_ptr, const State& state, TrackList& msckf_tracks, TrackList& slam_tracks,
                     std::vector<int>& anchor_idxs, TrackList& opp_tracks) {
    vio_updater_.getMsckfTracks(msckf_tracks);
    vio_updater_.getSlamTracks(slam_tracks, anchor_idxs, state.nPosesMax());
    opp_tracks = vio_updater_.trackManager().getOppTracks();
    state_ptr = std::make_shared<SimpleState>(state.getDynamicStates(), column(state.getPositionArray()),
                                              column(state.getOrientationArray()), column(state.getFeatureArray()),
                                              state.getCovariance(), anchor_idxs);
  }

  /** vio.cpp:498-574 with the correspondences delivered by the caller: the peer's snapshot becomes a SimpleState, the
   *  SLAM-SLAM correspondences are fused now (Ekf::processOthersMeasurement -> Updater::collaborativeUpdate: MultiSlamUpdate,
   *  pair fuseCI, applyCI), the MSCKF-MSCKF ones wait for the next visual update (vio_updater.cpp:185), as in the
   *  reference.  Returns nullopt when this agent has no SLAM track (vio.cpp:529-531) or no correspondence was found. */
  std::optional<State> processOtherMeasurements(double timestamp, const int uav_id, const Vectorx& dynamic_state,
                                                const Vectorx& positions_state, const Vectorx& orientations_state,
                                                const Vectorx& features_state, const Matrix& cov,
                                                const TrackListPtr& received_msckf_trcks_ptr,
                                                const std::vector<int>& anchor_idxs,
                                                const std::vector<Correspondence>& slam_correspondences,
                                                const std::vector<Correspondence>& msckf_correspondences) {
    std::shared_ptr<SimpleState> ptr = std::make_shared<SimpleState>(dynamic_state, positions_state, orientations_state,
                                                                     features_state, cov, anchor_idxs);
    TrackList current_slam;
    std::vector<int> current_anchors;
    vio_updater_.getSlamTracks(current_slam, current_anchors, params_.n_poses_max);
    if (current_slam.empty() && vio_updater_.trackManager().getOppTracks().empty()) return std::nullopt;
    if (slam_correspondences.empty() && msckf_correspondences.empty()) return std::nullopt;   // !place_found
    SlamMatches slam_matches;
    for (const auto& c : slam_correspondences)
      slam_matches.emplace_back(uav_id, static_cast<int>(c.current), c.received, ptr);
    MsckfMatches msckf_matches;
    for (const auto& c : msckf_correspondences) {
      if (c.received < 0 || static_cast<size_t>(c.received) >= received_msckf_trcks_ptr.size() ||
          !received_msckf_trcks_ptr[static_cast<size_t>(c.received)])
        throw std::invalid_argument("processOtherMeasurements: MSCKF correspondence outside the received track list");
      const TrackPtr& trk = received_msckf_trcks_ptr[static_cast<size_t>(c.received)];
      msckf_matches.emplace_back(uav_id, c.current, trk->getId(), trk, ptr);
    }
    vio_updater_.setSlamMatches(slam_matches);
    vio_updater_.setMsckfMatches(msckf_matches);
    return ekf_.processOthersMeasurement(timestamp);
  }
#endif

  /** The operator objects behind the facade (the reference keeps them private; the parity tests read them). */
  Ekf& ekf() { return ekf_; }
  VioUpdater& vioUpdater() { return vio_updater_; }
  [[nodiscard]] const Params& params() const { return params_; }

 private:
  /** vio.cpp:372-434: the 10-double match vector (cam_id, t_prev, x_prev, y_prev, t_curr, x_curr, y_curr, 3-D truth) ->
   *  MatchList with the measured (distorted) pixel coordinates; the undistortion of :399-405 is part of
   *  TrackManager::manageTracks here (xb_tm_manage_tracks). */
  MatchList importMatches(const std::vector<double>& match_vector, const unsigned int seq, TiledImage&) const {
    const unsigned int feature_arr_blk_sz = 10;
    if (match_vector.size() % feature_arr_blk_sz != 0) throw std::invalid_argument("match vector: 10 doubles per match");
    const unsigned int n_matches = static_cast<unsigned int>(match_vector.size() / feature_arr_blk_sz);
    MatchList matches(n_matches);
    for (unsigned int i = 0; i < n_matches; ++i) {
      const double* m = &match_vector[feature_arr_blk_sz * i];
      matches[i].previous = Feature(m[1], seq - 1, 0.0, 0.0, m[2], m[3], -1.0);
      matches[i].current = Feature(m[4], seq, 0.0, 0.0, m[5], m[6], -1.0);
    }
    return matches;
  }
  /** Eigen's Quaternion::setFromTwoVectors (the rotation that takes a to b along the shortest arc). */
  static Quaternion fromTwoVectors(const Vector3& a, const Vector3& b) {
    const Vector3 v0 = a.normalized(), v1 = b.normalized();
    const double c = v1.dot(v0);
    if (c < -1.0 + 1e-12) {   // antiparallel: any axis orthogonal to v0
      const Vector3 helper = std::fabs(v0(0)) < 0.9 ? Vector3(1, 0, 0) : Vector3(0, 1, 0);
      const Vector3 axis = v0.cross(helper).normalized();
      return Quaternion(0.0, axis(0), axis(1), axis(2));
    }
    const Vector3 axis = v0.cross(v1);
    const double s = std::sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
    return Quaternion(s * 0.5, axis(0) * invs, axis(1) * invs, axis(2) * invs);
  }
  static Vectorx column(const Matrix& m) {
    Vectorx v(m.rows());
    for (int i = 0; i < static_cast<int>(m.rows()); ++i) v(i) = m(i, 0);
    return v;
  }
  static std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  static std::string unquote(const std::string& s) {
    return s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front() ? s.substr(1, s.size() - 2) : s;
  }

  VioUpdater vio_updater_;   // declared before ekf_: Ekf keeps a reference to it (vio.h:225-236)
  Ekf ekf_;
  Params params_;
  double msckf_baseline_x_n_{0}, msckf_baseline_y_n_{0};
  Camera camera_;
  Tracker tracker_;
  TrackManager track_manager_;
  StateManager state_manager_;
  RangeMeasurement last_range_measurement_;
  SunAngleMeasurement last_angle_measurement_;
  Feature lrf_img_pt_n_;
  bool initialized_{false};
  bool self_init_start_{false}, initialize_start_{false};
  std::vector<Vector3> imu_data_batch_;
};

}  // namespace x
