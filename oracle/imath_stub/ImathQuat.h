/* TEST INFRASTRUCTURE ONLY: named by pspRT.h, nothing of it is used. */
