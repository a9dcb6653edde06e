"""Host logic of boxsegliver_b200/inference.py that needs no GPU: which mirrored passes run_TTA makes."""
from boxsegliver_b200.inference import tta_variants


def test_tta_variants_follow_reference_bit_tests():
    # entry/infer_2d.py:64-75: `flip & 1`, `flip & 2`, `flip & 3 > 0`
    assert tta_variants(0, True) == [0]
    assert tta_variants(3, False) == [0]
    assert tta_variants(1, True) == [0, 1, 3]
    assert tta_variants(2, True) == [0, 2, 3]
    assert tta_variants(3, True) == [0, 1, 2, 3]
    # entry/main_eval_3d.py:250-284
    assert tta_variants(7, True, three_d=True) == [0, 1, 2, 3, 4, 5, 6, 7]
    assert tta_variants(1, True, three_d=True) == [0, 1, 3, 5, 7]
