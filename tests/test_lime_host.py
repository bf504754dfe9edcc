"""Host side of the AudioLIME path (src/lime_explainer.py:283-301, 380-407): perturbation rows, cosine kernel and the weighted
ridge surrogate, checked against scikit-learn (the library the reference's LIME runs on)."""
import numpy as np
import pytest

from audio_deepfake_explainability_b200 import lime_explainer as le


def _fake_probs(masks, true_w=(0.30, -0.10, 0.05, 0.20), bias=0.35, seed=1):
    rng = np.random.default_rng(seed)
    p = bias + masks.astype(np.float64) @ np.asarray(true_w) + 0.01 * rng.standard_normal(len(masks))
    p = np.clip(p, 0.0, 1.0)
    return np.stack([1.0 - p, p], axis=1)


def test_masks_follow_the_lime_recipe():
    m = le.lime_masks(1000, 4, 0)
    ref = np.random.RandomState(0).randint(0, 2, 1000 * 4).reshape(1000, 4)
    ref[0, :] = 1
    assert m.dtype == np.uint8 and np.array_equal(m, ref)
    assert len(np.unique(m, axis=0)) == 16                                   # BASELINE configs[4]: only 16 distinct masks


def test_kernel_and_distances_match_sklearn():
    from sklearn.metrics import pairwise_distances
    m = le.lime_masks(200, 4, 3)
    d = le.cosine_distances_to_first(m)
    ref = pairwise_distances(m.astype(np.float64), m[0:1].astype(np.float64), metric="cosine").ravel()
    assert np.abs(d - ref).max() < 1e-12
    assert d[0] == 0.0 and np.all(d[(m == 0).all(1)] == 1.0)                 # an all-zero row is at distance 1
    k = le.lime_kernel(d, 0.25)
    assert np.abs(k - np.sqrt(np.exp(-(ref ** 2) / 0.25 ** 2))).max() < 1e-15


@pytest.mark.parametrize("seed", [0, 7])
def test_surrogate_matches_sklearn_ridge(seed):
    from sklearn.linear_model import Ridge
    masks = le.lime_masks(500, 4, seed)
    probs = _fake_probs(masks, seed=seed)
    exp = le.fit_lime(masks, probs)
    w = le.lime_kernel(le.cosine_distances_to_first(masks), 0.25)
    model = Ridge(alpha=1, fit_intercept=True).fit(masks.astype(np.float64), probs[:, exp.top_label], sample_weight=w)
    coef = np.array([dict(exp.local_exp)[i] for i in range(4)])
    assert np.abs(coef - model.coef_).max() < 1e-10
    assert abs(exp.intercept - model.intercept_) < 1e-10
    assert abs(exp.score - model.score(masks.astype(np.float64), probs[:, exp.top_label], sample_weight=w)) < 1e-10
    assert abs(exp.local_pred - model.predict(masks[0:1].astype(np.float64))[0]) < 1e-10
    # local_exp is sorted by |coef| descending (ties keep feature order), like sorted(zip(features, coef), key=abs, reverse=True)
    assert [i for i, _ in exp.local_exp] == sorted(range(4), key=lambda i: abs(model.coef_[i]), reverse=True)


def test_component_influences_reproduce_the_reference_pairing():
    masks = le.lime_masks(400, 4, 0)
    exp = le.fit_lime(masks, _fake_probs(masks))                             # top label = real (p < 0.5) or fake, either way
    names = le.COMPONENT_NAMES_4STEMS
    # reference (:403-407): zip(component names, local_exp) -> name k gets the k-th largest |weight|, whatever its feature id
    assert list(exp.component_influences) == list(names)
    assert list(exp.component_influences.values()) == [w for _, w in exp.local_exp]
    assert {n: exp.by_feature[n] for n in names} == {names[i]: w for i, w in exp.local_exp}
    mags = np.abs(list(exp.component_influences.values()))
    assert np.all(mags[:-1] >= mags[1:])


def test_top_label_and_shapes():
    masks = le.lime_masks(50, 4, 2)
    probs = _fake_probs(masks, bias=0.7)
    assert le.fit_lime(masks, probs).top_label == 1                          # unperturbed row predicts fake
    assert le.fit_lime(masks, 1.0 - probs).top_label == 0
    with pytest.raises(ValueError):
        le.fit_lime(masks, probs[:, :1])
    with pytest.raises(ValueError):
        le.fit_lime(masks, probs, component_names=("a", "b"))
    with pytest.raises(TypeError):
        le.explain_stems(np.zeros((4, 100), np.float32), predictor=object())
