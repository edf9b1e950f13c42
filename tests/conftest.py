import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def oracle_ipadic(oracle_mod):
    """IPADIC built by the oracle's restatement of the reference builder."""
    return oracle_mod.load_ipadic()


@pytest.fixture(scope="session")
def oracle_tok(oracle_mod, oracle_ipadic):
    return oracle_mod.OracleTokenizer(oracle_ipadic)


@pytest.fixture(scope="session")
def vocab(oracle_ipadic):
    from kanpyo_b200 import corpus
    return corpus.Vocabulary(oracle_ipadic.keywords, oracle_ipadic.morphs)


@pytest.fixture(scope="session")
def gpu_ipadic(oracle_ipadic):
    """The product Dict over the same arrays as the oracle's dictionary."""
    from helpers import to_product_dict
    return to_product_dict(oracle_ipadic)


@pytest.fixture(scope="session")
def gpu_tok(gpu_ipadic):
    import kanpyo_b200
    return kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
