"""AudioLIME stem explanation - drop-in for the perturb + predict + fit part of ``src/lime_explainer.py``.

The reference (``explain_predictions_separate``, :303-482) hands a Spleeter 4-stem factorisation to audioLIME's
``LimeAudioExplainer(kernel_width=0.25).explain_instance(factorization, predict_fn, num_samples, top_labels=1)`` and reads
``explanation.local_exp[top_label]`` (:380-407).  Both libraries are third-party and absent here (audioLIME @ CPJKU and
``lime``, versions unpinned, vendored by the author under the git-ignored ``XAIMethods/``); their published algorithm is
restated below, with the reference's own call site as the anchor:

  * ``LimeAudioExplainer.data_labels``: ``data = random_state.randint(0, 2, num_samples * n_features)`` reshaped to
    ``[num_samples, n_features]``, row 0 set to ones; every row is composed (sum of the selected stems) and predicted -
    in the reference one waveform at a time through ``predict_fn_unified`` (:283-301); here ALL rows go through one batched
    device sweep (``B200Predictor.stem_mask_sweep``), de-duplicated to the <= 2^n distinct masks.
  * distances = cosine distance of every row to row 0; kernel = ``sqrt(exp(-d^2 / kernel_width^2))``.
  * ``LimeBase.explain_instance_with_data`` with ``num_features=100000`` and ``feature_selection='auto'`` -> every feature
    is kept ('highest_weights' branch) and the surrogate is ``Ridge(alpha=1, fit_intercept=True)`` with the kernel as
    sample weights; ``local_exp[label]`` = ``(feature_id, coef)`` sorted by ``|coef|`` descending.
  * The reference then zips the component NAMES with that |coef|-sorted list (:403-407), i.e. names are paired by
    position, not by feature id.  ``component_influences`` reproduces that pairing (drop-in); ``by_feature`` is the
    id-correct mapping.

Spleeter separation itself (TensorFlow) stays outside: the stems are an input.
"""
from __future__ import annotations

from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from .sonics_api import B200Predictor

COMPONENT_NAMES_4STEMS = ("vocals0", "drums0", "bass0", "other0")      # SpleeterFactorization, spleeter:4stems, one temporal segment


class LimeExplanation(NamedTuple):
    top_label: int                                  # 1 = fake, 0 = real (argmax of the unperturbed prediction, top_labels=1)
    local_exp: List[Tuple[int, float]]              # (feature id, ridge coefficient), |coef| descending
    intercept: float
    score: float                                    # weighted R^2 of the surrogate
    local_pred: float                               # surrogate prediction at the all-ones row
    component_influences: Dict[str, float]          # the reference's name <-> weight pairing (by sorted position, :403-407)
    by_feature: Dict[str, float]                    # name -> coefficient of that stem
    masks: np.ndarray                               # [num_samples, n_features] uint8
    probabilities: np.ndarray                       # [num_samples, 2] = (1 - p, p)


def predict_fn_unified(waveforms: np.ndarray, predictor) -> np.ndarray:
    """``[N, 2] = (real_prob, fake_prob)`` of ``[N, samples]`` (or ``[samples]``) waveforms (:283-301) - one batched pass."""
    w = np.asarray(waveforms, dtype=np.float32)
    if w.ndim == 1:
        w = w[np.newaxis, :]
    if isinstance(predictor, B200Predictor):
        p = np.asarray(predictor.predict_batch(w), dtype=np.float64)
    else:
        p = np.array([float(predictor.predict(x, sr=44100)) for x in w], dtype=np.float64)
    return np.stack([1.0 - p, p], axis=1)


def lime_masks(num_samples: int, n_features: int, random_state=None) -> np.ndarray:
    """Perturbation rows of ``LimeAudioExplainer.data_labels``: Bernoulli(1/2) bits, row 0 all ones."""
    rs = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
    data = rs.randint(0, 2, num_samples * n_features).reshape((num_samples, n_features))
    data[0, :] = 1
    return data.astype(np.uint8)


def cosine_distances_to_first(data: np.ndarray) -> np.ndarray:
    """``sklearn.metrics.pairwise_distances(data, data[0:1], metric='cosine').ravel()``; an all-zero row has distance 1."""
    x = np.asarray(data, dtype=np.float64)
    ref = x[0]
    num = x @ ref
    den = np.sqrt((x * x).sum(1)) * np.sqrt((ref * ref).sum())
    cos = np.divide(num, den, out=np.zeros_like(num), where=den > 0)
    return np.clip(1.0 - cos, 0.0, 2.0)


def lime_kernel(distances: np.ndarray, kernel_width: float = 0.25) -> np.ndarray:
    d = np.asarray(distances, dtype=np.float64)
    return np.sqrt(np.exp(-(d ** 2) / kernel_width ** 2))


def weighted_ridge(X: np.ndarray, y: np.ndarray, w: np.ndarray, alpha: float = 1.0) -> Tuple[np.ndarray, float, float]:
    """``Ridge(alpha, fit_intercept=True).fit(X, y, sample_weight=w)`` in closed form: (coef, intercept, weighted R^2).
    The intercept is not penalised: centre with the weighted means, solve ``(Xc' W Xc + alpha I) b = Xc' W yc``."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    sw = w.sum()
    xm = (w[:, None] * X).sum(0) / sw
    ym = (w * y).sum() / sw
    Xc, yc = X - xm, y - ym
    A = Xc.T @ (w[:, None] * Xc) + alpha * np.eye(X.shape[1])
    coef = np.linalg.solve(A, Xc.T @ (w * yc))
    intercept = float(ym - xm @ coef)
    pred = X @ coef + intercept
    ss_res = float((w * (y - pred) ** 2).sum())
    ss_tot = float((w * (y - ym) ** 2).sum())
    score = 1.0 - ss_res / ss_tot if ss_tot > 0 else 0.0
    return coef, intercept, score


def fit_lime(masks: np.ndarray, probabilities: np.ndarray, component_names: Sequence[str] = COMPONENT_NAMES_4STEMS,
             kernel_width: float = 0.25, label: Optional[int] = None) -> LimeExplanation:
    """Surrogate fit of ``LimeBase.explain_instance_with_data`` on the perturbation rows and their ``[N, 2]`` predictions."""
    data = np.asarray(masks)
    probs = np.asarray(probabilities, dtype=np.float64)
    if data.ndim != 2 or probs.shape != (data.shape[0], 2):
        raise ValueError(f"masks [N, F] and probabilities [N, 2] expected, got {data.shape} / {probs.shape}")
    if len(component_names) != data.shape[1]:
        raise ValueError(f"{len(component_names)} component names for {data.shape[1]} features")
    top = int(np.argmax(probs[0])) if label is None else int(label)          # top_labels=1: the label of the unperturbed row
    weights = lime_kernel(cosine_distances_to_first(data), kernel_width)
    coef, intercept, score = weighted_ridge(data, probs[:, top], weights, alpha=1.0)
    order = sorted(range(data.shape[1]), key=lambda i: np.abs(coef[i]), reverse=True)     # stable, like sorted(zip(...))
    local_exp = [(int(i), float(coef[i])) for i in order]
    local_pred = float(data[0].astype(np.float64) @ coef + intercept)
    influences = {name: w for name, (_, w) in zip(component_names, local_exp)}            # the reference's pairing (:403-407)
    by_feature = {name: float(coef[i]) for i, name in enumerate(component_names)}
    return LimeExplanation(top, local_exp, intercept, score, local_pred, influences, by_feature, data.astype(np.uint8), probs)


def explain_stems(stems: np.ndarray, predictor: B200Predictor, num_samples: int = 500,
                  component_names: Sequence[str] = COMPONENT_NAMES_4STEMS, kernel_width: float = 0.25,
                  random_state=None, deduplicate: bool = True) -> LimeExplanation:
    """The hot part of ``explain_predictions_separate`` for one track whose stems are already separated.

    ``stems``: float ``[n_stems, L]`` (the components of the factorisation; their sum is the mix).  With 4 stems only 16
    distinct masks exist, so the device sweep evaluates each once (the reference recomputes duplicates) unless
    ``deduplicate=False``."""
    if not isinstance(predictor, B200Predictor):
        raise TypeError(f"explain_stems needs a B200Predictor, got {type(predictor).__name__}")
    stems = np.ascontiguousarray(np.asarray(stems, dtype=np.float32))
    masks = lime_masks(num_samples, stems.shape[0], random_state)
    if deduplicate:
        uniq, inverse = np.unique(masks, axis=0, return_inverse=True)
        probs = predictor.stem_mask_sweep(stems, uniq)[np.asarray(inverse).reshape(-1)]
    else:
        probs = predictor.stem_mask_sweep(stems, masks)
    return fit_lime(masks, probs, component_names, kernel_width)
