"""cherryml.utils helpers (reference cherryml/utils.py)."""
import os

from cherryml_b200 import utils


def test_get_process_args_is_the_modulo_striping():
    for rank, world, items in ((0, 3, list(range(10))), (2, 3, list("abcdefg")), (1, 1, [1, 2]), (5, 8, [1, 2]),
                               (0, 1, []), (3, 4, list(range(3)))):
        want = [items[i] for i in range(len(items)) if i % world == rank]  # reference utils.py:59-67
        assert utils.get_process_args(rank, world, items) == want


def test_pushd_restores_the_working_directory(tmp_path):
    before = os.getcwd()
    with utils.pushd(str(tmp_path)):
        assert os.path.realpath(os.getcwd()) == os.path.realpath(str(tmp_path))
    assert os.getcwd() == before
    try:
        with utils.pushd(str(tmp_path)):
            raise RuntimeError
    except RuntimeError:
        pass
    assert os.getcwd() == before


def test_get_families_and_amino_acids(tmp_path):
    for name in ("b.txt", "a.txt", "c.profiling", "d.e.txt"):
        (tmp_path / name).write_text("")
    assert utils.get_families(str(tmp_path)) == ["a", "b", "d"]
    assert utils.get_amino_acids() == utils.amino_acids and utils.get_amino_acids() is not utils.amino_acids
    assert len(utils.amino_acids) == 20
