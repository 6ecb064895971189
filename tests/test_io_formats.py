"""On-disk formats either side of the hot path: round trips and the reference's conventions
(notebooks/retrieval.ipynb cells 1 and 3; src/train.py:1143-1208, 3940-3984, 4009-4016)."""
import json

import numpy as np
import pytest

from patent_image_retrieval_b200 import io as pio


def test_gallery_cache_round_trip_and_reference_file_names(tmp_path):
    emb = np.random.default_rng(0).standard_normal((7, 16))            # float64 on disk -> float32 in memory
    paths = [f"/data/gallery/p{i}/fig_{i}.png" for i in range(7)]
    npy, js = pio.save_gallery_cache(tmp_path / "embeddings", "clip_vit", emb, paths)
    assert npy.name == "clip_vit.npy" and js.name == "clip_vit.json"
    assert json.load(open(js)) == paths and np.load(npy).shape == (7, 16)
    got, got_paths = pio.load_gallery_cache(tmp_path / "embeddings", "clip_vit")
    assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"] and got_paths == paths
    np.testing.assert_allclose(got, emb.astype(np.float32))
    assert pio.load_gallery_cache(tmp_path / "embeddings", "other_model") is None
    with pytest.raises(ValueError):
        pio.save_gallery_cache(tmp_path, "bad", emb, paths[:3])


def _training_data():
    rng = np.random.default_rng(1)
    offsets = {"patents": 0, "medium_cpcs": 40, "big_cpcs": 70, "main_cpcs": 82}
    return pio.TrainingData(
        X_figures=rng.standard_normal((30, 512)),                        # float64, as np.savez of a notebook array
        Y_pos=np.array([[0, 3], [1, 5], [1, 7], [99, 2], [2, 200]]), Y_neg=np.array([[0, 9], [1, 11]]),
        implication=np.array([[3, 45], [45, 72], [72, 85]]), exclusion=np.zeros((0, 2), dtype=np.int32),
        label_offsets=offsets, positive_figure_pairs=np.array([[0, 1], [1, 2], [5, 6], [7, 400]]),
        negative_figure_pairs=None)


def test_training_data_round_trip(tmp_path):
    td = _training_data()
    pio.save_training_data(tmp_path, td)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["label_offsets.json", "training_data.npz"]
    with np.load(tmp_path / "training_data.npz") as z:
        assert z["Y_pos"].dtype == np.int32 and "negative_figure_pairs" not in z.files
    got = pio.load_training_data(tmp_path)
    assert got.X_figures.dtype == np.float32 and got.X_figures.shape == (30, 512)
    np.testing.assert_array_equal(got.Y_pos, td.Y_pos)
    assert got.exclusion.shape == (0, 2) and got.negative_figure_pairs is None
    np.testing.assert_array_equal(got.positive_figure_pairs, td.positive_figure_pairs)
    assert got.label_offsets == td.label_offsets
    assert got.num_patents == 40 and got.label_num() == 40 + 30 + 12 + 9      # src/train.py:4009-4016
    (tmp_path / "label_offsets.json").write_text(json.dumps({"patents": 0}))
    with pytest.raises(KeyError):
        pio.load_training_data(tmp_path)


def test_figure_maps_follow_train_py():
    td = _training_data()
    f2p = pio.figure_to_pos_patent(td.Y_pos, num_figures=30, num_labels=td.label_num())
    assert f2p == {0: 3, 1: 7}                        # last pair wins; figure 99 and label 200 are out of range
    f2f = pio.figure_to_pos_figures(td.positive_figure_pairs, num_figures=30)
    assert f2f == {0: [1], 1: [0, 2], 2: [1], 5: [6], 6: [5]}          # symmetric; (7, 400) dropped
    assert pio.figure_to_pos_figures(None) == {}


def test_ground_truth_csr_matches_by_file_name(tmp_path):
    gallery = ["/g/a/x1.png", "/g/b/x2.png", "/g/c/x3.png", "/g/d/x4.png"]
    gt = {"q1.png": {"patent_positives": ["x3.png", "x1.png", "missing.png", "x1.png"], "cpc_positives": ["x2.png"]},
          "q3.png": {"patent_positives": []}}
    p = tmp_path / "ground_truth_2019.json"
    p.write_text(json.dumps(gt))
    keep, off, items, n_tot = pio.positives_csr(pio.load_ground_truth(p), ["/q/q1.png", "/q/q2.png", "q3.png"], gallery)
    assert keep == [0, 2]                             # q2 has no ground truth: skipped, as in the notebook
    assert off.tolist() == [0, 2, 2] and items.tolist() == [0, 2]
    assert n_tot.tolist() == [3, 0]                   # |P| counts the positive that is not in the gallery
    _, off2, items2, _ = pio.positives_csr(gt, ["q1.png"], gallery, key="cpc_positives")
    assert off2.tolist() == [0, 1] and items2.tolist() == [1]


def test_evaluation_results_have_the_notebook_keys(tmp_path):
    names = ["mrr", "ap", "ndcg"] + [f"{m}@{k}" for k in (5, 10, 20) for m in ("mrr", "precision", "recall")]
    per = np.arange(2 * len(names), dtype=np.float64).reshape(2, len(names)) / 100
    res = pio.save_evaluation_results(tmp_path / "results" / "evaluation_results_m.json", per, names)
    on_disk = json.load(open(tmp_path / "results" / "evaluation_results_m.json"))
    assert on_disk == res
    assert list(res) == ["query_wise_metrics", "summary_metrics"]
    assert list(res["summary_metrics"]) == ["MRR", "MRR@5", "MRR@20", "mAP", "mNDCG", "Recall@5", "Recall@10",
                                            "Recall@20", "Precision@5", "Precision@10", "Precision@20"]
    assert list(res["query_wise_metrics"]) == ["reciprocal_ranks", "reciprocal_ranks@5", "reciprocal_ranks@20",
                                               "ap_scores", "ndcg_scores", "recall_5", "recall_10", "recall_20",
                                               "precision_5", "precision_10", "precision_20"]
    assert res["summary_metrics"]["mAP"] == pytest.approx(per[:, 1].mean())
    assert res["query_wise_metrics"]["recall_10"] == pytest.approx(per[:, names.index("recall@10")].tolist())
    with pytest.raises(KeyError):
        pio.evaluation_results(per[:, :3], names[:3])
