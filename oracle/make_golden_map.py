"""Pins oracle/map_oracle.py against the UNMODIFIED reference scorer (src/predict.py voc_eval / voc_ap), CPU only:
writes the VOC xml annotations, the image list and the per-class detection files the reference reads into a temp
directory, calls the reference's own methods, and asserts rec / prec / ap equal the oracle's for every class, with the
VOC07 11-point metric and with the area metric.  Writes tests/golden/voc_map.npz (detections, ground truth, APs)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import map_oracle, ref_shim  # noqa: E402

CLASSES = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable', 'dog',
           'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor')


def main():
    ref_shim.load_reference()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import src.predict as predict
    n_img = 60
    gts = map_oracle.synthetic_ground_truth(n_img, seed=4)
    rng = np.random.RandomState(5)
    # synthetic kept boxes: jittered copies of the ground truth (true positives, duplicates) plus random boxes, with
    # deliberately repeated confidences (ties) and multi-class boxes as validation mode produces
    kept = []
    for i in range(n_img):
        boxes = []
        for (c, x1, y1, x2, y2, d) in gts[i]:
            for rep in range(rng.randint(0, 3)):
                j = rng.uniform(-0.03, 0.03, 4)
                cx, cy = (x1 + x2) / 2 / 416 + j[0], (y1 + y2) / 2 / 416 + j[1]
                w, h = (x2 - x1) / 416 * (1 + j[2]), (y2 - y1) / 416 * (1 + j[3])
                conf = np.float32(rng.choice([0.9, 0.75, 0.5, rng.uniform(0.1, 1.0)]))
                box = [np.float32(cx), np.float32(cy), np.float32(w), np.float32(h), conf, np.float32(rng.uniform(0.3, 1)), c]
                if rng.rand() < 0.3:
                    box += [np.float32(rng.uniform(0.01, 0.3)), int(rng.randint(0, 20))]
                boxes.append(box)
        for _ in range(rng.randint(0, 4)):
            box = [np.float32(v) for v in rng.uniform(0.1, 0.9, 2)] + [np.float32(v) for v in rng.uniform(0.05, 0.5, 2)]
            box += [np.float32(rng.uniform(0.01, 1)), np.float32(rng.uniform(0.05, 1)), int(rng.randint(0, 20))]
            boxes.append(box)
        kept.append(boxes)
    sizes = [(416, 416)] * n_img
    rows = map_oracle.detection_rows(kept, sizes)

    ev = object.__new__(predict.PASCALVOCEval)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, 'ann'))
        with open(os.path.join(tmp, 'test.txt'), 'w') as f:
            for i in range(n_img):
                f.write('%06d\n' % i)
        for i in range(n_img):
            objs = ''.join('<object><name>%s</name><pose>x</pose><truncated>0</truncated><difficult>%d</difficult>'
                           '<bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax><ymax>%d</ymax></bndbox></object>'
                           % (CLASSES[c], d, x1, y1, x2, y2) for (c, x1, y1, x2, y2, d) in gts[i])
            with open(os.path.join(tmp, 'ann', '%06d.xml' % i), 'w') as f:
                f.write('<annotation>%s</annotation>' % objs)
        # the reference's own file format (src/predict.py:172), written from the un-rounded values
        for c in range(20):
            with open(os.path.join(tmp, 'det_%s.txt' % CLASSES[c]), 'w') as f:
                for i, boxes in enumerate(kept):
                    for box in boxes:
                        x1 = (box[0] - box[2] / 2.0) * 416
                        y1 = (box[1] - box[3] / 2.0) * 416
                        x2 = (box[0] + box[2] / 2.0) * 416
                        y2 = (box[1] + box[3] / 2.0) * 416
                        for j in range(int((len(box) - 5) / 2)):
                            if int(box[6 + 2 * j]) == c:
                                f.write('%s %f %f %f %f %f\n' % ('%06d' % i, box[4] * box[5 + 2 * j], x1, y1, x2, y2))
        aps = {}
        orig_argsort = np.argsort
        np.argsort = lambda a, *args, **kw: orig_argsort(a, *args, **dict(kw, kind='stable'))  # define the tie order
        if not hasattr(np, 'bool'):
            np.bool = bool
        try:
            for metric07 in (True, False):
                for c in range(20):
                    with np.errstate(divide='ignore', invalid='ignore'):
                        rec_r, prec_r, ap_r = ev.voc_eval(os.path.join(tmp, 'det_{:s}.txt'), os.path.join(tmp, 'ann', '{:s}.xml'),
                                                          os.path.join(tmp, 'test.txt'), CLASSES[c],
                                                          os.path.join(tmp, 'cache'), 0.5, metric07)
                        rec_o, prec_o, ap_o = map_oracle.voc_eval(rows.get(c, []), gts, c, 0.5, metric07)
                    assert np.array_equal(rec_r, rec_o, equal_nan=True) and np.array_equal(prec_r, prec_o, equal_nan=True), c
                    assert ap_r == ap_o or (np.isnan(ap_r) and np.isnan(ap_o)), (c, ap_r, ap_o)
                    aps[(metric07, c)] = ap_o
        finally:
            np.argsort = orig_argsort
    flat = []
    for i, boxes in enumerate(kept):
        for box in boxes:
            for j in range(int((len(box) - 5) / 2)):
                flat.append([i, float(box[0]), float(box[1]), float(box[2]), float(box[3]), float(box[4]),
                             float(box[5 + 2 * j]), int(box[6 + 2 * j])])
    gt_flat = [[i, c, x1, y1, x2, y2, d] for i, objs in enumerate(gts) for (c, x1, y1, x2, y2, d) in objs]
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'voc_map.npz'), dets=np.array(flat, dtype=np.float32),
                        gts=np.array(gt_flat, dtype=np.int64), n_images=np.int64(n_img),
                        ap07=np.array([aps[(True, c)] for c in range(20)]), ap_area=np.array([aps[(False, c)] for c in range(20)]))
    with open(os.path.join(ROOT, 'tests', 'golden', 'PINNING.txt'), 'a') as f:
        f.write("voc_eval / voc_ap (60 synthetic images, ties, difficult flags, multi-class rows): oracle == reference on rec, prec, "
                "ap for all 20 classes, VOC07 and area metrics; mAP07 %.6f (oracle/make_golden_map.py)\n"
                % np.mean([aps[(True, c)] for c in range(20)]))
    print("pinned: voc_eval equal for 20 classes x 2 metrics; mAP07 = %.6f" % np.mean([aps[(True, c)] for c in range(20)]))


if __name__ == '__main__':
    main()
