"""NumPy restatement of the reference's PASCAL-VOC scorer (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows src/predict.py: the per-class detection rows written by the eval loop (:157-173: for every kept box and every
(cls_conf, cls_id) pair, prob = box_conf*cls_conf and the corners scaled by the image size, formatted with
'%s %f %f %f %f %f'), voc_ap (:239-263, VOC07 11-point or area under the monotone envelope) and voc_eval (:265-395:
detections sorted by -confidence, greedy matching against the image's ground truth with the +1 pixel convention,
'difficult' boxes ignored).  The text round trip of the reference ('%f' = 6 decimals, parsed back as float64) is part
of the arithmetic and is reproduced with the same formatting.  np.argsort is made stable (ties keep file order): the
reference's default quicksort leaves the order of equal confidences unspecified.
Pinned against the unmodified reference by oracle/make_golden_map.py (it writes the files the reference reads)."""
import numpy as np


def text_round(v):
    """float -> the float64 the reference reads back from its '%f' text files."""
    return float('%f' % v)


def detection_rows(kept_boxes_per_image, image_sizes):
    """src/predict.py:157-173.  kept_boxes_per_image: list (per image) of boxes [x,y,w,h,box_conf,(cls_conf,cls_id)+];
    returns dict cls_id -> list of (image index, prob, x1, y1, x2, y2) after the text round trip, in file order."""
    rows = {}
    for i, boxes in enumerate(kept_boxes_per_image):
        width, height = image_sizes[i]
        for box in boxes:
            x1 = (box[0] - box[2] / 2.0) * width
            y1 = (box[1] - box[3] / 2.0) * height
            x2 = (box[0] + box[2] / 2.0) * width
            y2 = (box[1] + box[3] / 2.0) * height
            box_conf = box[4]
            for j in range(int((len(box) - 5) / 2)):
                cls_conf = box[5 + 2 * j]
                cls_id = int(box[6 + 2 * j])
                prob = box_conf * cls_conf
                rows.setdefault(cls_id, []).append((i, text_round(prob), text_round(x1), text_round(y1), text_round(x2),
                                                    text_round(y2)))
    return rows


def voc_ap(rec, prec, use_07_metric=False):
    """src/predict.py:239-263."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            if np.sum(rec >= t) == 0:
                p = 0
            else:
                p = np.max(prec[rec >= t])
            ap = ap + p / 11.
    else:
        mrec = np.concatenate(([0.], rec, [1.]))
        mpre = np.concatenate(([0.], prec, [0.]))
        for i in range(mpre.size - 1, 0, -1):
            mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
        i = np.where(mrec[1:] != mrec[:-1])[0]
        ap = np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])
    return ap


def voc_eval(rows, gt_per_image, cls_id, ovthresh=0.5, use_07_metric=False):
    """src/predict.py:305-395 for one class.  rows: list of (image index, confidence, x1, y1, x2, y2) in file order;
    gt_per_image: list (per image) of (cls_id, xmin, ymin, xmax, ymax, difficult) integer tuples.
    Returns (rec, prec, ap)."""
    class_recs = {}
    npos = 0
    for i, objs in enumerate(gt_per_image):
        R = [o for o in objs if o[0] == cls_id]
        bbox = np.array([o[1:5] for o in R])
        difficult = np.array([o[5] for o in R]).astype(bool)
        npos = npos + int(np.sum(~difficult))
        class_recs[i] = {'bbox': bbox, 'difficult': difficult, 'det': [False] * len(R)}
    image_ids = [r[0] for r in rows]
    confidence = np.array([float(r[1]) for r in rows])
    BB = np.array([[float(z) for z in r[2:]] for r in rows]).reshape(-1, 4)
    sorted_ind = np.argsort(-confidence, kind='stable')
    BB = BB[sorted_ind, :]
    image_ids = [image_ids[x] for x in sorted_ind]
    nd = len(image_ids)
    tp = np.zeros(nd)
    fp = np.zeros(nd)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        bb = BB[d, :].astype(float)
        ovmax = -np.inf
        BBGT = R['bbox'].astype(float)
        if BBGT.size > 0:
            ixmin = np.maximum(BBGT[:, 0], bb[0])
            iymin = np.maximum(BBGT[:, 1], bb[1])
            ixmax = np.minimum(BBGT[:, 2], bb[2])
            iymax = np.minimum(BBGT[:, 3], bb[3])
            iw = np.maximum(ixmax - ixmin + 1., 0.)
            ih = np.maximum(iymax - iymin + 1., 0.)
            inters = iw * ih
            uni = ((bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) +
                   (BBGT[:, 2] - BBGT[:, 0] + 1.) * (BBGT[:, 3] - BBGT[:, 1] + 1.) - inters)
            overlaps = inters / uni
            ovmax = np.max(overlaps)
            jmax = np.argmax(overlaps)
        if ovmax > ovthresh:
            if not R['difficult'][jmax]:
                if not R['det'][jmax]:
                    tp[d] = 1.
                    R['det'][jmax] = 1
                else:
                    fp[d] = 1.
        else:
            fp[d] = 1.
    fp = np.cumsum(fp)
    tp = np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return rec, prec, ap


def mean_ap(rows_by_class, gt_per_image, num_classes=20, ovthresh=0.5, use_07_metric=True):
    """src/predict.py:397-437 (_do_python_eval): AP per class (classes without detections score on an empty file),
    mean over the classes."""
    aps = []
    for c in range(num_classes):
        with np.errstate(divide='ignore', invalid='ignore'):
            _, _, ap = voc_eval(rows_by_class.get(c, []), gt_per_image, c, ovthresh, use_07_metric)
        aps.append(ap)
    return aps, float(np.mean(aps))


def synthetic_ground_truth(n_images, seed=4, size=416, num_classes=20):
    """1-5 boxes per image, class U{0..19}, centre U[0.1,0.9], size U[0.05,0.5] (normalised), turned
    into the integer pixel corners a VOC xml holds; every 7th box is flagged 'difficult'."""
    rng = np.random.RandomState(seed)
    out = []
    k = 0
    for _ in range(n_images):
        objs = []
        for _ in range(rng.randint(1, 6)):
            c = int(rng.randint(0, num_classes))
            cx, cy = rng.uniform(0.1, 0.9, 2)
            w, h = rng.uniform(0.05, 0.5, 2)
            xmin = max(int((cx - w / 2) * size), 1)
            ymin = max(int((cy - h / 2) * size), 1)
            xmax = min(int((cx + w / 2) * size), size)
            ymax = min(int((cy + h / 2) * size), size)
            k += 1
            objs.append((c, xmin, ymin, xmax, ymax, 1 if k % 7 == 0 else 0))
        out.append(objs)
    return out
