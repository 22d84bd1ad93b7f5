package nnet;

/**
 * JNI binding of libfastnn.so (include/fastnn.h) for FastNeighborNet.  Not compiled in this repository (no JDK in the
 * build image); fastnn_jni.c next to this file is the native half.  Call sites: see INTEGRATION.md.
 */
final class NativeNN {
    static { System.loadLibrary("fastnn_jni"); }          // libfastnn_jni.so links libfastnn.so

    /** mode: NetMakerOriginal.NMMode.ordinal() - 0 Canonical, 1 Relaxed, 2 Random_N, 3 Random_NLOGN, 4 Random_LOGN. */
    static native int[] order(double[][] D, int nTaxa, int mode, int mult, boolean additive, long seed);

    /** Above 46 341 taxa the Java heap cannot index the packed triangle (DistancesAndNames.java:37,45): pass the file. */
    static native int[] orderFromFile(String phylipPath, int nTaxa, int mode, int mult, boolean additive, long seed);

    /** x[n(n-1)/2] in the live indexing of FastNN.java:409-418: entry (i, j) is the split {ordering[i+1..j]}. */
    static native double[] splitWeights(int[] ordering, double[] dUpper, int nTaxa);

    /**
     * The whole run of FastNN.main after option parsing: load the Phylip file, order, weigh, keep x > cutoff and print
     * the Nexus document to nexusPath (null = stdout).  Returns the circular ordering.
     */
    static native int[] network(String phylipPath, String nexusPath, int mode, int mult, boolean additive, long seed,
                                double cutoff, boolean printDistances);
}
