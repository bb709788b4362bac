"""FGD / feature-distance / diversity scores of the TED evaluation.

Mirror of the reference's ``EmbeddingSpaceEvaluator`` (scripts/model/ted_evaluator.py:13-154) as
``scripts/test_RAG_ted.py:35, 86, 129-140`` drives it: the constructor loads the autoencoder checkpoint
(``{'pose_dim', 'gen_dict'}``), ``push_samples(generated_poses, real_poses)`` encodes both pose batches to 32-d
features - on the device, through ``ls_pose_features`` (embedding_net.py in this package) - and keeps them as numpy
arrays; ``get_scores`` / ``get_diversity_scores`` are the reference's host-side numpy / scipy statistics over the
collected features (32 x 32 covariances, one matrix square root), restated here.
"""
import numpy as np
import torch
from scipy import linalg

from .embedding_net import EmbeddingNet

device = torch.device("cuda:0")


def _sqrtm(m):
    """scipy.linalg.sqrtm quietly: the reference passes disp=False (:124), an argument newer scipy releases dropped."""
    try:
        return linalg.sqrtm(m, disp=False)[0]
    except TypeError:
        return linalg.sqrtm(m)


class EmbeddingSpaceEvaluator:
    def __init__(self, embed_net_path="/p300/wangchy/zhiyh/ted_output/gesture_autoencoder_checkpoint_best.bin"):
        checkpoint = torch.load(embed_net_path, map_location="cpu")     # {'pose_dim': int, 'gen_dict': state_dict}
        self.pose_dim = int(checkpoint["pose_dim"])
        net = EmbeddingNet(self.pose_dim, 34)                            # 34-frame clips (:17)
        net.load_state_dict(checkpoint["gen_dict"])                      # strict: encoder AND decoder keys
        net.eval()
        net.freeze_pose_nets()
        self.net = net
        self.reset()

    def reset(self):
        self.real_feat_list, self.generated_feat_list, self.recon_err_diff = [], [], []

    def push_samples(self, generated_poses, real_poses):
        """Both [B, 34, pose_dim] on a CUDA device (:35-41)."""
        for poses, keep in ((real_poses, self.real_feat_list), (generated_poses, self.generated_feat_list)):
            feat = self.net(poses, variational_encoding=False)[0]        # the mu head (ls_pose_features)
            keep.append(feat.cpu().numpy())

    def get_features_for_viz(self):
        import umap                      # optional dependency of the reference (:48-57); not needed for the scores
        generated, real = np.vstack(self.generated_feat_list), np.vstack(self.real_feat_list)
        both = umap.UMAP().fit_transform(np.vstack((generated, real)))
        n = int(both.shape[0] / 2)
        return both[n:, :], both[0:n, :]

    def get_no_of_samples(self):
        return len(self.real_feat_list)

    def get_scores(self):
        """(frechet_dist, feat_dist) of generated vs real features (:59-88)."""
        generated, real = np.vstack(self.generated_feat_list), np.vstack(self.real_feat_list)
        try:
            frechet = self.calculate_frechet_distance(generated.mean(axis=0), np.cov(generated, rowvar=False),
                                                      real.mean(axis=0), np.cov(real, rowvar=False))
        except ValueError:
            frechet = float("inf")       # the reference's literal 1e+10000000000000
        feat_dist = np.mean([np.sum(np.absolute(r - g)) for r, g in zip(real, generated)])
        return frechet, feat_dist

    @staticmethod
    def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
        """|mu1 - mu2|^2 + tr(sigma1 + sigma2 - 2 sqrt(sigma1 sigma2)) (:90-145, after pytorch-fid)."""
        mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
        sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
        assert mu1.shape == mu2.shape, 'Training and test mean vectors have different lengths'
        assert sigma1.shape == sigma2.shape, 'Training and test covariances have different dimensions'
        root = _sqrtm(sigma1 @ sigma2)
        if not np.all(np.isfinite(root)):          # nearly singular product: regularise both covariances
            print('fid calculation produces singular product; adding %s to diagonal of cov estimates' % eps)
            jitter = eps * np.eye(len(sigma1))
            root = _sqrtm((sigma1 + jitter) @ (sigma2 + jitter))
        if np.iscomplexobj(root):                  # numerical noise may leave a small imaginary part
            if not np.allclose(np.diagonal(root).imag, 0, atol=1e-3):
                raise ValueError('Imaginary component {}'.format(np.max(np.abs(root.imag))))
            root = np.real(root)
        d = mu1 - mu2
        return float(d @ d) + np.trace(sigma1) + np.trace(sigma2) - 2.0 * np.trace(root)

    def get_diversity_scores(self):
        """Mean L1 distance between the first 500 pushed feature batches and 500 randomly drawn ones (:147-154)."""
        batches = self.generated_feat_list
        head = np.vstack(batches[:500])
        order = torch.randperm(len(batches))[:500]               # the one draw of the global generator (:149)
        shuffled = np.vstack([batches[int(k)] for k in order])
        return np.abs(head - shuffled).sum(axis=-1).mean()
