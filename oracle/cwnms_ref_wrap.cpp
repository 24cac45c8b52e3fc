// TEST INFRASTRUCTURE.  extern "C" doors into the reference's own post-processing (compiled from
// /root/reference/examples/YOLO-Master-Cross-Platform-Edge-Deployment/cpp/src/common.cpp, see oracle/Makefile):
//   cwnms_ref_decode_nms   = yolomaster::decode (decode_candidates :93-125 + nms_and_cap :127-205) on one raw (4+nc, A) prediction
//   cwnms_ref_nms_and_cap  = yolomaster::nms_and_cap on explicit candidates
// Used by tests/golden/make_cwnms_golden.py and tests/test_nms_oracle.py to pin oracle/nms_oracle.cw_nms.  Never on the product path.
#include "yolomaster.hpp"

#include <cstring>

using namespace yolomaster;

static Config make_cfg(int nc, float conf, float iou, int max_det, int cluster, float sigma) {
    Config cfg;
    cfg.conf_thresh = conf;
    cfg.iou_thresh = iou;
    cfg.max_det = max_det;
    cfg.multi_label = false;
    cfg.nms_mode = cluster ? NmsMode::ClusterWeighted : NmsMode::Standard;
    cfg.cw_sigma = sigma;
    cfg.class_names.assign(nc, "c");
    return cfg;
}

static int emit(const std::vector<Detection>& dets, float* out6, int cap) {
    int n = 0;
    for (const auto& d : dets) {
        if (n >= cap) break;
        float* o = out6 + 6 * n++;
        o[0] = d.box.x; o[1] = d.box.y; o[2] = d.box.width; o[3] = d.box.height; o[4] = d.conf; o[5] = static_cast<float>(d.class_id);
    }
    return n;
}

extern "C" {

// pred: (4 + nc, A) row-major fp32, xywh (centre) in letterboxed pixels; identity letterbox with the given frame size.
int cwnms_ref_decode_nms(const float* pred, int nc, int A, float conf, float iou, int max_det, int cluster, float sigma,
                         int frame_w, int frame_h, float* out6, int cap) {
    Config cfg = make_cfg(nc, conf, iou, max_det, cluster, sigma);
    LetterboxInfo lb;
    lb.orig_w = frame_w;
    lb.orig_h = frame_h;
    return emit(decode(pred, 4 + nc, A, cfg, lb), out6, cap);
}

// candidates: boxes (n,4) top-left xywh fp32, scores (n), classes (n)
int cwnms_ref_nms_and_cap(const float* boxes, const float* scores, const int* classes, int n, float conf, float iou, int max_det,
                          int cluster, float sigma, int frame_w, int frame_h, float* out6, int cap) {
    Config cfg = make_cfg(0, conf, iou, max_det, cluster, sigma);
    std::vector<RawDet> cands(n);
    for (int i = 0; i < n; ++i) {
        cands[i].box = cv::Rect2f(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]);
        cands[i].score = scores[i];
        cands[i].cls = classes[i];
    }
    return emit(nms_and_cap(cands, cfg, frame_w, frame_h), out6, cap);
}

}  // extern "C"
