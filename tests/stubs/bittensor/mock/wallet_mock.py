def get_mock_wallet(*a, **k):
    raise NotImplementedError("the bittensor stub has no wallets")
