run base --args "--steps 20" A=1
run serial RMCV_SERIAL=1
