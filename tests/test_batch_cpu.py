"""Host logic of the batch drivers that needs no GPU: token lists, batching, argument errors raised before any device work."""
import pytest

from magphase_b200 import batch


def test_read_tokens_and_batches(tmp_path):
    scp = tmp_path / 'file_id.scp'
    scp.write_text('# list of tokens\nhvd_593\n\n  hvd_594  # trailing comment\n#hvd_595\nhvd_596\n')
    assert batch.read_tokens(str(scp)) == ['hvd_593', 'hvd_594', 'hvd_596']
    assert list(batch._batches(list(range(7)), 3)) == [[0, 1, 2], [3, 4, 5], [6]]
    assert list(batch._batches([], 3)) == []


def test_argument_errors_come_before_device_work(tmp_path):
    with pytest.raises(ValueError):
        batch.run_waveform_generation(['a'], str(tmp_path), str(tmp_path / 'o'), 60, 45, 48000, pf_type='bogus')
    # an empty token list is a no-op on both sides
    assert batch.run_feature_extraction([], str(tmp_path), str(tmp_path / 'f'))['utterances'] == 0
    assert batch.run_waveform_generation([], str(tmp_path), str(tmp_path / 'w'), 60, 45, 48000)['frames'] == 0


def test_cli_parser():
    with pytest.raises(SystemExit):
        batch.main(['extract'])                      # missing required arguments
