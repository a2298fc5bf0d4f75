// mock
