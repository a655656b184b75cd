"""Chain output: dataset paths, shapes and the thinning rule of mcmc_sampler_output (src/mcmc_io_util.cpp:521-990)."""
import numpy as np
import pytest

from gw_analysis_tools_b200 import chain_io


def test_data_dump_layout(tmp_path):
    rng = np.random.default_rng(0)
    n, steps, dim = 3, 40, 5
    pos = rng.normal(size=(n, steps, dim))
    ll = rng.normal(size=(n, steps, 2))
    ids, temps = [0, 8, 16], [1.0, 1.0, 1.0]
    ac = rng.integers(1, 9, size=(n, dim))
    path = tmp_path / "dump.gwd"
    chain_io.write_data_dump(path, ids, temps, pos, ll, trim_lengths=[3, 0, 5], ac_values=ac)
    d = chain_io.read_dump(path)
    want = {"/MCMC_OUTPUT/CHAIN %d" % i for i in ids} | {"/MCMC_OUTPUT/LOGL_LOGP/CHAIN %d" % i for i in ids} | {
        "/MCMC_METADATA/CHAIN TEMPERATURES", "/MCMC_METADATA/SUGGESTED TRIM LENGTHS", "/MCMC_METADATA/AC VALUES"}
    assert set(d) == want
    for k, i in enumerate(ids):
        assert np.array_equal(d["/MCMC_OUTPUT/CHAIN %d" % i], pos[k]) and d["/MCMC_OUTPUT/CHAIN %d" % i].shape == (steps, dim)
        assert np.array_equal(d["/MCMC_OUTPUT/LOGL_LOGP/CHAIN %d" % i], ll[k])
    assert np.array_equal(d["/MCMC_METADATA/CHAIN TEMPERATURES"], temps)
    assert d["/MCMC_METADATA/SUGGESTED TRIM LENGTHS"].dtype == np.int32 and list(d["/MCMC_METADATA/SUGGESTED TRIM LENGTHS"]) == [3, 0, 5]
    assert np.array_equal(d["/MCMC_METADATA/AC VALUES"], ac)
    # without likelihoods and autocorrelation lengths those datasets are absent, trim lengths default to zero
    chain_io.write_data_dump(path, ids, temps, pos)
    d = chain_io.read_dump(path)
    assert not any("LOGL_LOGP" in k or "AC VALUES" in k for k in d) and list(d["/MCMC_METADATA/SUGGESTED TRIM LENGTHS"]) == [0, 0, 0]


def _thin_reference(pos, ac, trim):
    """write_flat_thin_output + count_indep_samples restated (src/mcmc_io_util.cpp:521-552, 555-580)."""
    n, steps, dim = pos.shape
    max_acs = [max(1, int(ac[i].max())) for i in range(n)]
    mean_ac = sum(max_acs) / n
    mean_pos = sum(steps - trim[i] for i in range(n)) / n
    indep = int(mean_pos / mean_ac)
    rows = []
    for i in range(n):
        for j in range(trim[i], steps):
            if j % max_acs[i] == 0 and len(rows) < indep:
                rows.append(pos[i, j])
    return np.array(rows).reshape(-1, dim)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flat_thin_output_rule(tmp_path, seed):
    rng = np.random.default_rng(seed)
    n, steps, dim = 4, 200, 3
    pos = rng.normal(size=(n, steps, dim))
    ac = rng.integers(1, 12, size=(n, dim))
    trim = list(rng.integers(0, 40, size=n))
    path = tmp_path / "thin.gwd"
    rows = chain_io.write_flat_thin_output(path, pos, ac, trim)
    got = chain_io.read_dump(path)["/THINNED_MCMC_OUTPUT/THINNED FLATTENED CHAINS"]
    want = _thin_reference(pos, ac, trim)
    assert rows == want.shape[0] == got.shape[0]
    assert np.array_equal(got, want)


def test_bad_arguments(tmp_path):
    with pytest.raises(Exception):
        chain_io.write_flat_thin_output(tmp_path / "no_such_dir" / "x.gwd", np.zeros((1, 4, 2)), np.ones((1, 2)))
