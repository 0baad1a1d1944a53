"""Dataset formats of the reference's HeadNeRF training script (NeRFs/HeadNeRF/train/audio_exp_nerf.py:45-196, class GetData), written by
data_util/process_data.py:253-288:

    <data_dir>/transforms_exp_{train,val}.json   {"focal_len", "cx", "cy", "frames": [{"img_id", "aud_id", "transform_matrix" 4x4,
                                                  "face_rect" [x, y, w, h], "exp" [76 or 79]}]}
    <data_dir>/<aud_file> (aud.npy)               (T, 16, 29) DeepSpeech windows
    <data_dir>/bc.jpg                             background, RGB
    <data_dir>/<gt_dirs>/<img_id>.jpg             target frames (the reference reads them with cv2: BGR channel order)
    <data_dir>/ori_imgs/<img_id>.lms              68 x 2 landmarks; rows 48.. are the mouth
    <data_dir>/parsing/<img_id>.png               face parsing; pure red = torso

Host side only (json / numpy / PIL), except that the N_rand training rays are generated on the device for the SELECTED pixels
(ops.get_rays_at) instead of building the 450 x 450 ray grid per sample and indexing it (:123-139,189-191)."""
import json
import os

import numpy as np
import torch

from . import ops


def _imread_rgb(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


class HeadDataset(torch.utils.data.Dataset):
    """GetData (audio_exp_nerf.py:45).  __getitem__ returns the reference's tuple
    (batch_rays (2, N_rand, 3), target_s (N_rand, 3), bc_rgb, auds (T, 16, 29), raw_img, pose (3, 4), exp, index) with device tensors."""

    def __init__(self, data_dir, aud_file, mode, args, skip=1, device="cuda"):
        self.data_dir, self.aud_file, self.mode, self.args, self.device = data_dir, aud_file, mode, args, torch.device(device)
        with open(os.path.join(data_dir, f"transforms_exp_{mode}.json")) as fp:
            self.meta = json.load(fp)
        self.aud_features = np.load(os.path.join(data_dir, aud_file))
        self.background_img = torch.tensor(_imread_rgb(os.path.join(data_dir, "bc.jpg")) / 255.0).to(self.device)
        self.focal, self.cx, self.cy = float(self.meta["focal_len"]), float(self.meta["cx"]), float(self.meta["cy"])
        self.H, self.W = int(self.cy * 2), int(self.cx * 2)
        self.skip = 1 if mode == "train" else getattr(args, "testskip", 1)
        self.all_imgs, self.all_parse_imgs, self.all_landmarks = [], [], []
        self.all_poses, self.all_face_rects, self.all_exprs, auds = [], [], [], []
        for frame in self.meta["frames"][::skip]:
            iid = str(frame["img_id"])
            self.all_imgs.append(os.path.join(data_dir, args.gt_dirs, iid + ".jpg"))
            self.all_landmarks.append(os.path.join(data_dir, "ori_imgs", iid + ".lms"))
            self.all_parse_imgs.append(os.path.join(data_dir, "parsing", iid + ".png"))
            self.all_poses.append(np.array(frame["transform_matrix"]))
            auds.append(self.aud_features[min(frame["aud_id"], self.aud_features.shape[0] - 1)])
            self.all_face_rects.append(np.array(frame["face_rect"], dtype=np.int32))
            self.all_exprs.append(frame["exp"])
        self.data_size = len(self.all_imgs)
        self.auds = torch.tensor(np.asarray(auds), dtype=torch.float).to(self.device)

    def __len__(self):
        return self.data_size

    def sample_pixels(self, face_rect, landmark, parse_img):
        """The pixel selection of sample_rays (:141-187), same region tests and the same order of np.random.choice draws
        (mouth, torso, face rectangle, outside), returned as (N_rand, 2) int64 (row, col) in the reference's concatenation order
        (rect, norect, mouth, torso).  NB the reference compares the ROW coordinate with the x ranges (coords[:, 0] is the row)."""
        a, H, W = self.args, self.H, self.W
        lm = landmark[48:]
        max_x, min_x = np.max(lm[:, 0]) + 20, np.min(lm[:, 0]) - 20
        max_y, min_y = np.max(lm[:, 1]) + 20, np.min(lm[:, 1]) - 20
        rr, cc = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        coords = np.stack([rr, cc], -1).reshape(-1, 2)
        mouth = (coords[:, 0] >= min_x) & (coords[:, 0] <= max_x) & (coords[:, 1] >= min_y) & (coords[:, 1] <= max_y)
        rect = (coords[:, 0] >= face_rect[0]) & (coords[:, 0] <= face_rect[0] + face_rect[2]) & \
               (coords[:, 1] >= face_rect[1]) & (coords[:, 1] <= face_rect[1] + face_rect[3])
        torso = ((parse_img[:, :, 0] == 255) & (parse_img[:, :, 1] == 0) & (parse_img[:, :, 2] == 0)).reshape(-1)
        c_mouth, c_rect, c_norect, c_torso = coords[mouth], coords[rect & ~mouth], coords[~rect], coords[torso]
        mouth_num, torso_num = a.mouth_rays, a.torso_rays
        sample_num = a.N_rand - mouth_num - torso_num
        rect_num = int(sample_num * a.sample_rate)
        norect_num = sample_num - rect_num
        s_mouth = c_mouth[np.random.choice(c_mouth.shape[0], size=[mouth_num], replace=False)]
        s_torso = c_torso[np.random.choice(c_torso.shape[0], size=[torso_num], replace=False)]
        s_rect = c_rect[np.random.choice(c_rect.shape[0], size=[rect_num], replace=False)]
        s_norect = c_norect[np.random.choice(c_norect.shape[0], size=[norect_num], replace=False)]
        return np.concatenate([s_rect, s_norect, s_mouth, s_torso], 0).astype(np.int64)

    def __getitem__(self, index):
        if index is None:
            index = np.random.choice(self.data_size)
        raw_img = torch.tensor(_imread_rgb(self.all_imgs[index])[:, :, ::-1].copy())            # cv2.imread order: BGR
        self.H, self.W = raw_img.shape[0], raw_img.shape[1]
        target = raw_img.to(self.device).float() / 255.0
        parse = _imread_rgb(self.all_parse_imgs[index])
        pose = self.all_poses[index][:3, :4]
        sel = torch.from_numpy(self.sample_pixels(self.all_face_rects[index], np.loadtxt(self.all_landmarks[index]), parse)).to(self.device)
        rays = ops.get_rays_at(sel, self.focal, torch.tensor(pose, dtype=torch.float32, device=self.device), 0.0, 1.0, self.cx, self.cy)
        batch_rays = torch.stack([rays[:, 0:3], rays[:, 3:6]], 0)
        target_s = target[sel[:, 0], sel[:, 1]]
        bc_rgb = self.background_img[sel[:, 0], sel[:, 1]] if self.mode == "train" else self.background_img
        exp = torch.tensor(self.all_exprs[index], dtype=torch.float32)
        return batch_rays, target_s, bc_rgb, self.auds, raw_img, pose, exp, index
