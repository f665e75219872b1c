"""B200-native latent-space clustering hot path (DEC layer + full-covariance GMM EM).

Drop-in for the clustering path of Julia310/Spectrogram-Cube-Clustering:

    from spectrogram_cube_clustering_b200.networks import ClusteringLayer, DEC
    from spectrogram_cube_clustering_b200.models import target_distribution, gmm, kmeans, batch_eval
    from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer

All N-sized work runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/scc_b200.h`` (``libscc_b200.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"
