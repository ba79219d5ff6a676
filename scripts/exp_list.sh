run base A=1
run serial RMCV_SERIAL=1
