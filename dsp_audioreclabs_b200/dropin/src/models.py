"""Drop-in for the 'knn' branch of src/models.py (:21-35,52-72,226-246): create_classifier('knn')
returns an object with fit / predict / evaluate whose neighbours come from the CUDA KNN
(exact float64 ranking, ties to the lower train index, vote ties to the smallest label --
the semantics of sklearn's KNeighborsClassifier the reference wraps).

The other classifier types are outside this hot path: they are
delegated unchanged to the reference's own src/models.py, loaded by file from the reference `src/`
directory this package found on sys.path (src/__init__.py), or to the module named by
DSP_REFERENCE_MODELS; without either they raise."""
import importlib
import importlib.util
import os
import sys

import numpy as np

from dsp_audioreclabs_b200 import batch as _b


def _metrics(y_true, y_pred):
    """accuracy / per-class report / confusion matrix with the keys the callers read
    (models.py:63-72; run_experiments.py:286-300 reads 'accuracy', 'confusion_matrix')."""
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    labels = np.unique(np.concatenate([y_true, y_pred]))
    idx = {l: i for i, l in enumerate(labels)}
    cm = np.zeros((len(labels), len(labels)), dtype=np.int64)
    for t, p in zip(y_true, y_pred):
        cm[idx[t], idx[p]] += 1
    report = {}
    tot = cm.sum()
    for l, i in idx.items():
        tp = cm[i, i]
        pp, sp = cm[:, i].sum(), cm[i].sum()
        prec = tp / pp if pp else 0.0
        rec = tp / sp if sp else 0.0
        f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
        report[str(l)] = {'precision': float(prec), 'recall': float(rec), 'f1-score': float(f1), 'support': float(sp)}
    acc = float(np.trace(cm) / tot) if tot else 0.0
    rows = [v for v in report.values()]
    sup = np.array([r['support'] for r in rows])
    report['accuracy'] = acc
    for name, w in (('macro avg', np.ones(len(rows)) / max(len(rows), 1)), ('weighted avg', sup / sup.sum() if sup.sum() else sup)):
        report[name] = {k: float(sum(r[k] * wi for r, wi in zip(rows, w))) for k in ('precision', 'recall', 'f1-score')}
        report[name]['support'] = float(sup.sum())
    return acc, report, cm


class TraditionalClassifier:
    """models.py:18-72, 'knn' only."""

    def __init__(self, classifier_type='knn', **kwargs):
        self.classifier_type = classifier_type
        if classifier_type != 'knn':
            raise ValueError(f"不支持的分类器类型: {classifier_type}")
        self.model = _b.KNN(n_neighbors=kwargs.get('n_neighbors', 3))

    def fit(self, X_train, y_train):
        self.model.fit(X_train, y_train)

    def predict(self, X_test):
        return self.model.predict(X_test)

    def evaluate(self, X_test, y_test):
        y_pred = self.predict(X_test)
        accuracy, report, cm = _metrics(y_test, y_pred)
        return {'accuracy': accuracy, 'predictions': y_pred, 'classification_report': report,
                'confusion_matrix': cm}


_ref_models = None


def _reference_models():
    global _ref_models
    if _ref_models is not None:
        return _ref_models
    name = os.environ.get('DSP_REFERENCE_MODELS')
    if name:
        try:
            _ref_models = importlib.import_module(name)
            return _ref_models
        except ImportError:
            pass
    import src as _pkg
    for d in getattr(_pkg, 'REFERENCE_SRC_DIRS', []):
        path = os.path.join(d, 'models.py')
        if os.path.isfile(path):
            spec = importlib.util.spec_from_file_location('src._reference_models', path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules['src._reference_models'] = mod
            spec.loader.exec_module(mod)
            _ref_models = mod
            return mod
    try:
        _ref_models = importlib.import_module('reference_models')
    except ImportError:
        _ref_models = None
    return _ref_models


def create_classifier(classifier_type, **kwargs):
    """models.py:226-246."""
    if classifier_type == 'knn':
        return TraditionalClassifier('knn', **kwargs)
    if classifier_type in ('naive_bayes', 'decision_tree', 'svm', 'mlp'):
        ref = _reference_models()
        if ref is None:
            raise ValueError(f"classifier '{classifier_type}' is outside the CUDA hot path and the reference's "
                             "src/models.py was not found (run through `python -m dsp_audioreclabs_b200.run`, or "
                             "set DSP_REFERENCE_ROOT / DSP_REFERENCE_MODELS)")
        return ref.create_classifier(classifier_type, **kwargs)
    raise ValueError(f"不支持的分类器类型: {classifier_type}")
