"""Scripted stand-ins for everything `fit_accumulate_test` (base.py:301-455) calls, shared by
the fixture generator (which drives the REFERENCE's own driver code) and by
tests/test_fit_driver.py (which drives the re-hosted driver).  Each scenario scripts a cost
trajectory that exercises a different exit of the loop."""
import numpy as np

K, D = 3, 2
LEN_VEC = [[7, 0, 7, 0, 0, 0, 0, 0, 1, 21], [5, 7, 12, 0, 0, 0, 0, 1, 1, 22]]
N = 12

SCENARIOS = {
    # name: (m_iter, threshold, cost trajectory as a function of the iteration)
    "converges": (40, 1e-3, lambda it: 5.0 + 3.0 * np.exp(-0.9 * it)),
    "runs_out": (6, 1e-9, lambda it: 5.0 + 1.0 / (1 + it) + 0.3 * (it % 2)),
    "stalls_after_best": (80, 1e-12, lambda it: 4.0 + 0.05 * it + (3.0 if it < 4 else 0.0) + 0.01 * np.sin(it)),
}


class ScriptedModel(object):
    """Mixin: deterministic _init/_predict_posteriors/_do_mstep; the driver under test comes
    from the class it is mixed into."""

    def script(self, name):
        self.traj = SCENARIOS[name][2]
        self.iteration = 0
        self.calls_this_iter = 0
        self.n_components, self.n_features = K, D
        self.tol, self.n_iter, self.verbose = 1e-7, 100, False
        self.finalized_with = None

    def _init(self, X, lengths=None):
        self.params_vec1 = np.arange(K * 4, dtype=np.float64).reshape(K, 4) / 10.0
        self.labels = np.zeros(N)
        self.labels_local = np.zeros(N)

    def _check(self):
        pass

    def _predict_posteriors(self, X, len_vec, region_id, m_queue):
        it = self.iteration
        rng = np.random.default_rng(1000 * it + region_id)
        n = len_vec[region_id][0]
        stats = {'post': rng.random(K), 'obs': rng.random((K, D)), 'obs*obs.T': rng.random((K, D, D))}
        labels = rng.integers(0, K, size=n)
        base = float(self.traj(it))
        c_pair = base * (0.3 + 0.1 * region_id)
        c_un = base * (0.7 - 0.1 * region_id)
        m_queue.put((region_id, stats, labels, 0.11 * base, c_pair, c_un, c_pair + c_un))
        self.calls_this_iter += 1
        if self.calls_this_iter == len(len_vec):
            self.calls_this_iter = 0
            self.iteration += 1
        return True

    def _initialize_sufficient_statistics(self):
        return {'nobs': 0, 'start': np.zeros(K), 'trans': np.zeros((K, K)), 'post': np.zeros(K),
                'obs': np.zeros((K, D)), 'obs**2': np.zeros((K, D)), 'obs*obs.T': np.zeros((K, D, D))}

    def _accumulate_sufficient_statistics_1(self, stats, stats1):
        stats['post'] += stats1['post']
        stats['obs'] += stats1['obs']
        stats['obs*obs.T'] += stats1['obs*obs.T']
        return stats

    def _do_mstep(self, stats):
        self.params_vec1 = self.params_vec1 + 0.01 * stats['post'].sum() + 0.001 * stats['obs*obs.T'].sum()

    def _ou_param_varied_constraint(self, params_vec):
        self.finalized_with = np.array(params_vec)

    def _sync_model(self):
        pass
