"""Guide-keyframe predictor of the body model: placeholder replaced by the KV-cached sampler."""


def load_guide_predictor(cp_path: str):
    raise NotImplementedError("guide-transformer keyframes (--resume_trans) are not built yet")
