run base --args "--steps 20" A=1
run nocopy --args "--steps 20" RMCV_EXP_NOCOPY=1
run serial RMCV_SERIAL=1
run serial_nocopy RMCV_SERIAL=1 RMCV_EXP_NOCOPY=1
